O=gpurun_out/r05f; mkdir -p $O
timeout 300 python tools/spec_debug.py > $O/spec_debug.log 2>&1; tail -30 $O/spec_debug.log | cut -c1-300
