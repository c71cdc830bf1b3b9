O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -12 $O/pytest.log
for f in "--no-fuse" ""; do for sl in 1 2; do
timeout 300 python tools/quick_bench.py --nt 438 --reps 4 --spinup 0 --land-slabs $sl --no-river $f > $O/quick438_s$sl$f.log 2>&1; echo "slabs $sl $f"; grep "^rep [23]" $O/quick438_s$sl$f.log | cut -c1-130
done; done
for sl in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --land-slabs $sl > $O/bench_s$sl.json 2> $O/bench_s$sl.err; python -c "
import json; d=json.load(open('$O/bench_s$sl.json')); k=d['roofline']['kernels']; print('slabs $sl', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'zone', round(k['zone_kernel']['ms'],2), round(k['zone_kernel']['ms_alone'],2), 'elem', round(k['element_kernel']['ms'],2), round(k['element_kernel']['ms_alone'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
