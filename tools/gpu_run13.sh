O=gpurun_out/r02o; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_synth.py -x -q -k "zone_terms_added or where_an_element_sits or slabs_and_time" > $O/pytest_fuse.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_fuse.log
for f in "--no-fuse" ""; do for sl in 1 2; do
timeout 300 python tools/quick_bench.py --nt 438 --reps 4 --spinup 0 --land-slabs $sl --no-river $f > $O/quick438_s$sl$f.log 2>&1; echo "slabs $sl $f"; grep "^rep [23]" $O/quick438_s$sl$f.log | cut -c1-130
done; done
