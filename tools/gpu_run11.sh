O=gpurun_out/r02m; mkdir -p $O
for v in _base ""; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python tools/quick_bench.py --nt 438 --reps 4 --spinup 0 --land-slabs 1 --no-river > $O/quick438$v.log 2>&1; grep "^rep [0123]" $O/quick438$v.log | cut -c1-200
done
ncu --set full --clock-control none --import-source on -k regex:'element_kernel' -s 2 -c 1 -o $O/elem_new python tools/quick_bench.py --nt 438 --reps 1 --spinup 2 --land-slabs 1 --no-river > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
