O=gpurun_out/r03r; mkdir -p $O
for wb in 1 2 3; do for sl in 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --river-plain-min 4000 --wave-blocks $wb > $O/bench_wb$wb.json 2> $O/bench_wb$wb.err; python -c "
import json; d=json.load(open('$O/bench_wb$wb.json')); k=d['roofline']['kernels']; print('wb $wb', round(d['ms_per_step'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), [round(v['ms_alone'],2) for n,v in k.items() if n.startswith('routing')], 'e2e', round(d['e2e']['ms_per_step'],2))"
done; done
