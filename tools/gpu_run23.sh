O=gpurun_out/r02y; mkdir -p $O
for wb in 1 2 3 6; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wave-blocks $wb > $O/bench_wb${wb}.json 2> $O/bench_wb${wb}.err; python -c "
import json; d=json.load(open('$O/bench_wb${wb}.json')); k=d['roofline']['kernels']; print('wave blocks $wb', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'zone', round(k['zone_kernel']['ms'],2), round(k['zone_kernel']['ms_alone'],2), 'elem', round(k['element_kernel']['ms'],2), round(k['element_kernel']['ms_alone'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), [round(v['ms_alone'],2) for n,v in k.items() if n.startswith('routing')], 'e2e', round(d['e2e']['ms_per_step'],2))"
done
