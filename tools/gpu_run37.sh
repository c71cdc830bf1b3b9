O=gpurun_out/r03l; mkdir -p $O
python tools/quick_bench.py --nt 438 --reps 3 --spinup 1 --land-slabs 1 --no-river > $O/quick.log 2>&1; grep "^rep [12]" $O/quick.log | cut -c1-120
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$i.json 2> $O/bench$i.err; python -c "
import json; d=json.load(open('$O/bench$i.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"; done
timeout 900 python tests/bench_configs.py --cases config2 > $O/bc.jsonl 2> $O/bc.err
python - <<'PY'
import json
for l in open('gpurun_out/r03l/bc.jsonl'):
    d=json.loads(l); print(d['case'], 'land', round(d['land_ms'],2), 'zone', round(d['zone_ms'],2), 'elem', round(d['element_ms'],2), 'river', d['river_ms'] and round(d['river_ms'],2), 'value %.3e'%d['value'])
PY
timeout 600 python -m pytest tests/test_gpu_synth.py tests/test_gpu_golden.py -x -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
