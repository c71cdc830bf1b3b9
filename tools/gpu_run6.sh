O=gpurun_out/r02f; mkdir -p $O
python -m pytest tests/test_gpu_synth.py -x -q -k "sits or members or slabs or ragged" > $O/pytest_place.log 2>&1; tail -4 $O/pytest_place.log
for n in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_worker.py > $O/multi_n$n.log 2>&1; echo "n=$n rc=$?"; grep -E "MULTI_GPU" $O/multi_n$n.log
done
