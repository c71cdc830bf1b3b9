O=gpurun_out/r03v; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_synth.py -x -q -k "pipelined or slabs_and_time or full_size" > $O/pytest1.log 2>&1; tail -2 $O/pytest1.log
for i in 1 2 3 4; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench$i.err; python -c "
import json; d=json.load(open('$O/bench$i.json')); print(round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'gauged', round(d['e2e_gauged_nodes']['ms_per_step'],2))"; done
