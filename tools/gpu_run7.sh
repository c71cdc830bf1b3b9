O=gpurun_out/r02h; mkdir -p $O
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-verify "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; python -c "
import json; d=json.load(open('$O/$name.json')); print(d['value'], d['ms_per_step'], d['per_rank_ms_per_step']['rows'])"; }
run chunk110 --land-chunk 110
run chunk73 --land-chunk 73
run chunk44 --land-chunk 44
NCCL_P2P_USE_CUDA_MEMCPY=1 run memcpy
