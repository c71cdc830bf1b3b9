O=gpurun_out/r03h; mkdir -p $O
for v in ""; do for sl in 1 2; do
echo "== lib$v slabs $sl"
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so timeout 900 python tests/bench_configs.py --cases config2,hland_96p,config5 --land-slabs $sl > $O/bc.jsonl 2> $O/bc.err
python - <<'PY'
import json
for l in open('gpurun_out/r03h/bc.jsonl'):
    d=json.loads(l); print(d['case'], 'land', round(d['land_ms'],2), 'zone', round(d['zone_ms'],2), 'elem', round(d['element_ms'],2), 'river', d['river_ms'] and round(d['river_ms'],2), 'value %.3e'%d['value'])
PY
tail -2 $O/bc.err
done; done

timeout 600 python -m pytest tests/test_gpu_synth.py tests/test_gpu_golden.py -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
