O=gpurun_out/r04b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_synth.py -x -q -k "small_zone_counts" > $O/pytest1.log 2>&1; tail -8 $O/pytest1.log
