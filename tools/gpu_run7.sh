O=gpurun_out/r02h; mkdir -p $O
run() { n=$1; shift; name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; python -c "
import json; d=json.load(open('$O/$name.json')); print(d['ms_per_step'], d['per_rank_ms_per_step']['rows'], d['multi_gpu_parity'])"; }
run 2 weak_n2_default --steps 20 --warmup 5 --no-verify
NCCL_MAX_NCHANNELS=1 run 2 weak_n2_ch1 --steps 20 --warmup 5 --no-verify
NCCL_MAX_NCHANNELS=1 NCCL_NTHREADS=64 run 2 weak_n2_ch1_t64 --steps 20 --warmup 5 --no-verify
run 2 weak_n2_verify --steps 10 --warmup 3
