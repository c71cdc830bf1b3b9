set -x
O=gpurun_out/r03o; mkdir -p $O
run() { n=$1; shift; name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; python -c "
import json; d=json.load(open('$O/$name.json')); print(round(d['value']/1e10,3), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e10,3), (d.get('e2e_gauged_nodes') or {}).get('value'), d.get('multi_gpu_parity'))"; }
run 8 bench_weak_n8 --steps 20 --warmup 5 --no-cpu-baseline
run 8 bench_strong_n8 --steps 20 --warmup 5 --scaling strong --no-cpu-baseline
run 8 bench_config1_n8 --config 1 --steps 20 --warmup 5 --no-cpu-baseline
run 8 bench_config4_n8 --config 4 --steps 5 --warmup 2 --no-cpu-baseline
run 4 bench_weak_n4 --steps 20 --warmup 5 --no-cpu-baseline
run 4 bench_strong_n4 --steps 20 --warmup 5 --scaling strong --no-cpu-baseline
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; tail -2 $O/pytest_multi.log
