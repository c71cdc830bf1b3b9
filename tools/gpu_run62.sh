O=gpurun_out/r05c; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "finite_smax or snow_redistribution or random_ragged or snow_classes" > $O/pytest_spec.log 2>&1; echo "rc=$?" >> $O/pytest_spec.log; tail -30 $O/pytest_spec.log
