O=gpurun_out/r05l; mkdir -p $O
timeout 600 python tests/bench_configs.py > $O/bench_configs.jsonl 2> $O/bench_configs.err; echo rc=$?
python - <<'PY'
import json
for l in open('gpurun_out/r05l/bench_configs.jsonl'):
    d=json.loads(l); print(d['case'], 'land', round(d['land_ms'],2), 'river', d['river_ms'] and round(d['river_ms'],2), 'value %.3e'%d['value'], 'dev', d['oracle_sample_max_rel_deviation'])
PY
python bench.py --config 2 > $O/bench_config2_n1.json 2> $O/bench_config2_n1.err; echo "config2 rc=$?"; python -c "
import json; d=json.load(open('$O/bench_config2_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'])"
