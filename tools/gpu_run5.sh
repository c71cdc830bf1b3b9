set -x
O=gpurun_out/r02e; mkdir -p $O
run() { n=$1; shift; name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; grep -E "\[bench\]|Error|error" $O/$name.err | tail -4; }
run 8 bench_weak_n8 --steps 20 --warmup 5
run 8 bench_strong_n8 --steps 20 --warmup 5 --scaling strong
run 8 bench_config1_n8 --config 1 --steps 20 --warmup 5
run 8 bench_config4_n8 --config 4 --steps 5 --warmup 2
run 4 bench_weak_n4 --steps 20 --warmup 5
run 4 bench_strong_n4 --steps 20 --warmup 5 --scaling strong
run 2 bench_weak_n2 --steps 20 --warmup 5
run 2 bench_strong_n2 --steps 20 --warmup 5 --scaling strong
python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; tail -3 $O/pytest_multi.log
