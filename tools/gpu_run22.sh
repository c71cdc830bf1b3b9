O=gpurun_out/r02x; mkdir -p $O
HB_DEBUG_TIMELINE=1 HB_DEBUG_MARKS=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_tl.json 2> $O/bench_tl.err
grep "hb timeline\|hb marks" $O/bench_tl.err | head -150 > $O/timeline.txt; wc -l $O/timeline.txt
