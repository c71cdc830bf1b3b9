O=gpurun_out/r05m; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or finite_smax or abi or errors" > $O/pytest_subset.log 2>&1; echo "rc=$?" >> $O/pytest_subset.log; tail -3 $O/pytest_subset.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err; python -c "
import json; d=json.load(open('$O/bench_short.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
