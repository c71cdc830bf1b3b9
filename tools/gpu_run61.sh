O=gpurun_out/r05b; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/bench_n1.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], r['kernel'], r['frac'], r['frac_alone'], d['gpu_launches'], d['clocks'])"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-300 $O/bench_reference.json
python bench.py --steps 200 --warmup 3 --no-cpu-baseline > $O/bench_10years_n1.json 2> $O/bench_10years_n1.err; echo "10y rc=$?"; python -c "
import json; d=json.load(open('$O/bench_10years_n1.json')); print('10 years:', d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('simulated_days_timed'))"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/b_short.json 2> $O/b_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/quick_bench.py --nt 438 --reps 2 --spinup 6 --land-slabs 1 > $O/quick438.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'zone_kernel|element_kernel|reach_wave_kernel' -s 18 -c 3 -o $O/prof_wet438 python tools/quick_bench.py --nt 438 --reps 2 --spinup 6 --land-slabs 1 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
