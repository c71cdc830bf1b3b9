O=gpurun_out/r04h; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "sibling or golden or hland_96c or 96c" > $O/pytest1.log 2>&1; tail -3 $O/pytest1.log
timeout 900 python tests/bench_configs.py --cases hland_96c,hland_96p > $O/bc.jsonl 2> $O/bc.err
python - <<'PY'
import json
for l in open('gpurun_out/r04h/bc.jsonl'):
    d=json.loads(l); print(d['case'], 'land', round(d['land_ms'],2), 'zone', round(d['zone_ms'],2), 'elem', round(d['element_ms'],2), 'river', d['river_ms'] and round(d['river_ms'],2), 'value %.3e'%d['value'], 'dev', d['oracle_sample_max_rel_deviation'])
PY
