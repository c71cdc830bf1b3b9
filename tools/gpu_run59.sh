O=gpurun_out/r04j; mkdir -p $O
for v in _u2 _u4; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python tools/quick_bench.py --nt 438 --reps 3 --spinup 1 --land-slabs 1 --no-river > $O/quick$v.log 2>&1; echo "lib$v"; grep "^rep [2]" $O/quick$v.log | cut -c1-110
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$v.json 2> $O/bench$v.err; python -c "
import json; d=json.load(open('$O/bench$v.json')); print('   bench', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
