O=gpurun_out/r02k; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -5 $O/pytest.log
for v in _base ""; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python tools/quick_bench.py --nt 438 --reps 2 --spinup 6 > $O/quick438$v.log 2>&1; grep "^rep [01]" $O/quick438$v.log | cut -c1-420
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e']['ms_per_step'])"
