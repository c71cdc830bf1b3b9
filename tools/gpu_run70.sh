O=gpurun_out/r05k; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log; tail -5 $O/pytest_multi.log | cut -c1-300
