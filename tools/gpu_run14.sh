O=gpurun_out/r02p; mkdir -p $O
timeout 300 python tools/quick_bench.py --nt 438 --reps 4 --spinup 0 --land-slabs 1 --no-river > $O/quick438_s1.log 2>&1; grep "^rep [23]" $O/quick438_s1.log | cut -c1-130
ncu --set full --clock-control none --import-source on -k regex:'element_kernel|zone_kernel' -s 4 -c 2 -o $O/fused python tools/quick_bench.py --nt 438 --reps 1 --spinup 2 --land-slabs 1 --no-river > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
