O=gpurun_out/r05a; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "ret_io" > $O/pytest_retio.log 2>&1; echo "rc=$?" >> $O/pytest_retio.log; tail -5 $O/pytest_retio.log
