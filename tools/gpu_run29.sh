O=gpurun_out/r03e; mkdir -p $O
for ch in 219 146 110 73; do for pm in 2048; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --land-chunk $ch > $O/bench_ch${ch}.json 2> $O/bench_ch${ch}.err; python -c "
import json; d=json.load(open('$O/bench_ch${ch}.json')); k=d['roofline']['kernels']; print('chunk $ch', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), [round(v['ms_alone'],2) for n,v in k.items() if n.startswith('routing')], 'e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])"
done; done
HB_DEBUG_TIMELINE=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_tl.json 2> $O/bench_tl.err
grep "hb timeline" $O/bench_tl.err | head -24
