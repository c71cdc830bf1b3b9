O=gpurun_out/r04e; mkdir -p $O
python tools/quick_bench.py --nt 438 --reps 3 --spinup 1 --land-slabs 1 --no-river > $O/quick.log 2>&1; grep "^rep [2]" $O/quick.log | cut -c1-330
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench$i.err; python -c "
import json; d=json.load(open('$O/bench$i.json')); print('   bench', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['config']['recstep_substeps_with_pow_frac'])"; done
timeout 600 python -m pytest tests/test_gpu_synth.py tests/test_gpu_golden.py -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
