O=gpurun_out/r03q; mkdir -p $O
for pm in 1000 2048 4000 8000 12000; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --river-plain-min $pm > $O/bench_pm$pm.json 2> $O/bench_pm$pm.err; python -c "
import json; d=json.load(open('$O/bench_pm$pm.json')); k=d['roofline']['kernels']; print('plain_min $pm', round(d['ms_per_step'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), [round(v['ms_alone'],2) for n,v in k.items() if n.startswith('routing')], 'e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])"
done
