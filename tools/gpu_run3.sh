set -x
O=gpurun_out/r02c; mkdir -p $O
python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; tail -5 $O/pytest_multi.log
for sc in weak strong; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --scaling $sc > $O/bench_n2_$sc.json 2> $O/bench_n2_$sc.err; echo "rc=$?"; tail -3 $O/bench_n2_$sc.err
done
