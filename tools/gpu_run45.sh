O=gpurun_out/r03t; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_synth.py -x -q -k "snow_classes" > $O/pytest1.log 2>&1; tail -5 $O/pytest1.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
