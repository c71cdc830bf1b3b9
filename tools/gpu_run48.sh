O=gpurun_out/r03w; mkdir -p $O
for i in 1 2 3 4 5 6; do HB_DEBUG_MARKS=1 HB_DEBUG_TIMELINE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench$i.err; python -c "
import json; d=json.load(open('$O/bench$i.json')); print($i, round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"; done
