O=gpurun_out/r02w; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_synth.py -x -q -k "slabs_and_time or zone_terms_added or pipelined_windows or full_size" > $O/pytest1.log 2>&1; tail -3 $O/pytest1.log
for i in 1 2 3; do python tools/quick_bench.py --nt 438 --reps 3 --spinup 1 > $O/quick$i.log 2>&1; grep "^rep [12]" $O/quick$i.log | cut -c1-150; done
for sl in 2 3 4; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --land-slabs $sl > $O/bench_s${sl}.json 2> $O/bench_s${sl}.err; python -c "
import json; d=json.load(open('$O/bench_s${sl}.json')); k=d['roofline']['kernels']; print('slabs $sl', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'zone', round(k['zone_kernel']['ms'],2), round(k['zone_kernel']['ms_alone'],2), 'elem', round(k['element_kernel']['ms'],2), round(k['element_kernel']['ms_alone'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), [round(v['ms_alone'],2) for n,v in k.items() if n.startswith('routing')], 'e2e', round(d['e2e']['ms_per_step'],2))"
done
