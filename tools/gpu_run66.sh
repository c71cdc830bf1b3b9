O=gpurun_out/r05g; mkdir -p $O
timeout 300 python tools/spec_debug2.py 746 > $O/dbg746.log 2>&1; cat $O/dbg746.log | cut -c1-260
timeout 300 python tools/spec_debug2.py 120 > $O/dbg120.log 2>&1; grep "nt=\|element" $O/dbg120.log | cut -c1-260
