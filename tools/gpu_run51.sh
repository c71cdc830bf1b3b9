O=gpurun_out/r04a; mkdir -p $O
for e in 0 1; do for i in 1 2; do
if [ $e = 1 ]; then export HB_EXP_WAIT_SLABS=1; else unset HB_EXP_WAIT_SLABS; fi
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_e${e}_$i.json 2> $O/bench_e${e}_$i.err; python -c "
import json; d=json.load(open('$O/bench_e${e}_$i.json')); print('exp $e', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done; done
