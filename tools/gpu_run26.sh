O=gpurun_out/r04i; mkdir -p $O
run() { n=$1; shift; name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; python -c "
import json; d=json.load(open('$O/$name.json')); print(round(d['value']/1e10,3), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e10,3), d.get('multi_gpu_parity'), d['per_rank_ms_per_step']['rows'])"; }
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi.log 2>&1; tail -3 $O/pytest_multi.log
run 2 bench_weak_n2 --steps 20 --warmup 5 --no-cpu-baseline
run 2 bench_strong_n2 --steps 20 --warmup 5 --scaling strong --no-cpu-baseline
