O=gpurun_out/r03d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_synth.py -x -q -k "plain_routing or full_size or musk_mct or long_reaches or ragged_problem_all" > $O/pytest1.log 2>&1; tail -4 $O/pytest1.log
python tools/quick_bench.py --nt 438 --reps 3 --spinup 1 > $O/quick.log 2>&1; grep "^rep [12]" $O/quick.log | cut -c1-150
for wb in 2 3 4; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wave-blocks $wb > $O/bench_wb${wb}.json 2> $O/bench_wb${wb}.err; python -c "
import json; d=json.load(open('$O/bench_wb${wb}.json')); k=d['roofline']['kernels']; print('wave blocks $wb', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), [round(v['ms_alone'],2) for n,v in k.items() if n.startswith('routing')], 'e2e', round(d['e2e']['ms_per_step'],2), 'launches', d['gpu_launches'])"
done
