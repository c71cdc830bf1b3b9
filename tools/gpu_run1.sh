set -x
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
python bench.py --steps 20 --warmup 5 > $O/bench_wet.json 2> $O/bench_wet.err; echo "rc=$?"
python bench.py --steps 20 --warmup 5 --forcing dry --window 240 --no-cpu-baseline > $O/bench_dry240.json 2> $O/bench_dry240.err
python bench.py --steps 20 --warmup 5 --window 240 --no-cpu-baseline > $O/bench_wet240.json 2> $O/bench_wet240.err
for wb in 1 2 3 6; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --wave-blocks $wb > $O/bench_wet_wb$wb.json 2> $O/bench_wet_wb$wb.err; done
python tools/quick_bench.py --nt 438 --reps 2 --spinup 8 > $O/quick438.log 2>&1
python tools/quick_bench.py --nt 240 --reps 2 --spinup 14 > $O/quick240.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'zone_kernel|element_kernel|reach_wave_kernel' -s 45 -c 3 -o $O/prof_wet240 python tools/quick_bench.py --nt 240 --reps 2 --spinup 14 > $O/ncu_full.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/b_short.json 2> $O/b_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
ls -la $O
