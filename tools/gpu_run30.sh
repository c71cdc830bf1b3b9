O=gpurun_out/r03f; mkdir -p $O
timeout 900 python tests/bench_configs.py --cases config2,config3,config5,hland_96p,hland_96c > $O/bench_configs.jsonl 2> $O/bench_configs.err; echo rc=$?
python - <<'PY'
import json
for l in open('gpurun_out/r03f/bench_configs.jsonl'):
    d=json.loads(l); print(d['case'], 'land', round(d['land_ms'],2), 'zone', round(d['zone_ms'],2), 'elem', round(d['element_ms'],2), 'river', d['river_ms'] and round(d['river_ms'],2), 'value %.3e'%d['value'], 'dev', d['oracle_sample_max_rel_deviation'])
PY
tail -3 $O/bench_configs.err
