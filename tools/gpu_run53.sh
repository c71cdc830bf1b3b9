O=gpurun_out/r04c; mkdir -p $O
for v in "" _zp64 _zp128; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python tools/quick_bench.py --nt 438 --reps 3 --spinup 1 --land-slabs 1 --no-river > $O/quick$v.log 2>&1; echo "lib$v"; grep "^rep [2]" $O/quick$v.log | cut -c1-110
for i in 1 2; do HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$v$i.json 2> $O/bench$v$i.err; python -c "
import json; d=json.load(open('$O/bench$v$i.json')); print('   bench', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; done
done
