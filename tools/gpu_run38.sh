O=gpurun_out/r03m; mkdir -p $O
for i in 1 2 3; do for v in "" _noeo; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$v$i.json 2> $O/bench$v$i.err; python -c "
import json; d=json.load(open('$O/bench$v$i.json')); print('lib$v', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks'])"; done; done
