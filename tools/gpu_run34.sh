O=gpurun_out/r03j; mkdir -p $O
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200_nouh.so ncu --set full --clock-control none --import-source on -k regex:'element_kernel' -s 3 -c 1 -o $O/lahn_elem python tests/bench_configs.py --cases config2 --land-slabs 1 --reps 1 > $O/ncu.log 2>&1
tail -2 $O/ncu.log
