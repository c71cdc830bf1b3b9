O=gpurun_out/r03g; mkdir -p $O
for opt in "--land-slabs 1" "--land-slabs 2" "--land-slabs 1 --no-land-fuse" "--land-slabs 2 --no-land-fuse"; do
echo "== $opt"
timeout 900 python tests/bench_configs.py --cases config2,config5,hland_96c $opt > $O/bc.jsonl 2> $O/bc.err
python - <<'PY'
import json
for l in open('gpurun_out/r03g/bc.jsonl'):
    d=json.loads(l); print(d['case'], 'land', round(d['land_ms'],2), 'zone', round(d['zone_ms'],2), 'elem', round(d['element_ms'],2), 'river', d['river_ms'] and round(d['river_ms'],2), 'value %.3e'%d['value'])
PY
done
