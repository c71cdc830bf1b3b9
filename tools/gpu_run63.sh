O=gpurun_out/r05d; mkdir -p $O
timeout 800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -40 $O/pytest_gpu.log | cut -c1-300
