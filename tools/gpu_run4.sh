set -x
O=gpurun_out/r02d; mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/bench_default_n1.json 2> $O/bench_default_n1.err; echo "rc=$?"; tail -3 $O/bench_default_n1.err
python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_config1_n1.json 2> $O/bench_config1_n1.err; echo "rc=$?"; tail -5 $O/bench_config1_n1.err
python bench.py --config 4 --members 16 --steps 5 --warmup 2 --no-cpu-baseline > $O/bench_config4_m16_n1.json 2> $O/bench_config4_m16_n1.err; echo "rc=$?"; tail -5 $O/bench_config4_m16_n1.err
/usr/bin/time -v python bench.py --config 4 --members 64 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_config4_m64_n1.json 2> $O/bench_config4_m64_n1.err; echo "rc=$?"; tail -25 $O/bench_config4_m64_n1.err | grep -E "Maximum resident|Elapsed|Error|error"
