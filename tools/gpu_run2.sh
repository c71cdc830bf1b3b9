set -x
O=gpurun_out/r02b; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -15 $O/pytest.log
for cfg in "1 4 0" "2 4 0" "2 2 0" "2 4 73" "2 2 73" "3 2 73" "4 2 55" "2 3 110" "2 6 73"; do set -- $cfg
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --land-slabs $1 --wave-blocks $2 --land-chunk $3 > $O/bench_s$1_w$2_c$3.json 2> $O/bench_s$1_w$2_c$3.err
done
ls -la $O
