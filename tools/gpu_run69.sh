O=gpurun_out/r05j; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], r['kernel'], r['frac'], r['frac_alone'], d['gpu_launches'], d['clocks'])"
timeout 600 python tests/bench_configs.py --cases smax_high,smax_high_general,smax_low,smax_low_general > $O/bc_smax.jsonl 2> $O/bc_smax.err; echo rc=$?
python - <<'PY'
import json
for l in open('gpurun_out/r05j/bc_smax.jsonl'):
    d=json.loads(l); print(d['case'], 'land', round(d['land_ms'],2), d['land_ms_reps'], 'zone', round(d['zone_ms'],2), 'elem', round(d['element_ms'],2), 'general', round(d['general_ms'],2), 'reruns/window', d['spec_reruns_per_window'], 'value %.3e'%d['value'])
PY
