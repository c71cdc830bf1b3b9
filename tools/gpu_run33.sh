O=gpurun_out/r03i; mkdir -p $O
for v in _base _nofast _nouh ""; do
echo "== lib$v"
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so timeout 900 python tests/bench_configs.py --cases config2 --land-slabs 1 > $O/bc.jsonl 2> $O/bc.err
python - <<'PY'
import json
for l in open('gpurun_out/r03i/bc.jsonl'):
    d=json.loads(l); print(d['case'], 'land', round(d['land_ms'],2), 'zone', round(d['zone_ms'],2), 'elem', round(d['element_ms'],2), 'river', d['river_ms'] and round(d['river_ms'],2), 'value %.3e'%d['value'])
PY
tail -1 $O/bc.err
done
