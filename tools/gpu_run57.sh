O=gpurun_out/r04g; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], r['kernel'], r['frac'], r['frac_alone'], d['gpu_launches'], d['clocks'])"
