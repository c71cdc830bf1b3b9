O=gpurun_out/r05h; mkdir -p $O
timeout 600 python tests/bench_configs.py --cases smax_high,smax_low,smax_general > $O/bc_smax.jsonl 2> $O/bc_smax.err; echo rc=$?
python - <<'PY'
import json
for l in open('gpurun_out/r05h/bc_smax.jsonl'):
    d=json.loads(l); print(d['case'], 'land', round(d['land_ms'],2), 'zone', round(d['zone_ms'],2), 'elem', round(d['element_ms'],2), 'general', round(d['general_ms'],2), 'n_spec', d['n_spec'], 'reruns/window', d['spec_reruns_per_window'], 'value %.3e'%d['value'], 'dev', d['oracle_sample_max_rel_deviation'])
PY
tail -3 $O/bc_smax.err
