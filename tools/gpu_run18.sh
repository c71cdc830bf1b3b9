O=gpurun_out/r02t; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_synth.py -x -q -k "out_of_range" > $O/pytest1.log 2>&1; tail -3 $O/pytest1.log
for sw in persistent grouped; do for sl in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --land-slabs $sl --river-sweep $sw > $O/bench_s${sl}_$sw.json 2> $O/bench_s${sl}_$sw.err; python -c "
import json; d=json.load(open('$O/bench_s${sl}_$sw.json')); k=d['roofline']['kernels']; print('slabs $sl $sw', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'zone', round(k['zone_kernel']['ms'],2), round(k['zone_kernel']['ms_alone'],2), 'elem', round(k['element_kernel']['ms'],2), round(k['element_kernel']['ms_alone'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), [round(v['ms_alone'],2) for n,v in k.items() if n.startswith('routing')], 'e2e', round(d['e2e']['ms_per_step'],2))"
done; done
