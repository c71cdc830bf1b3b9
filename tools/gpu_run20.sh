O=gpurun_out/r02v; mkdir -p $O
for v in _p1t1 _p2t1 _p1t2 ""; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python tools/quick_bench.py --nt 438 --reps 3 --spinup 1 > $O/quick$v.log 2>&1; echo "lib$v"; grep "^rep [12]" $O/quick$v.log | cut -c1-150
done
