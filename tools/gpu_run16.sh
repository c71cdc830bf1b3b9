O=gpurun_out/r02r; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_synth.py -x -q -k "out_of_range" > $O/pytest1.log 2>&1; tail -3 $O/pytest1.log
ncu --set full --clock-control none --import-source on -k regex:'element_kernel|zone_kernel' -s 4 -c 2 -o $O/fused python tools/quick_bench.py --nt 438 --reps 1 --spinup 2 --land-slabs 1 --no-river > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:'zone_kernel' -s 2 -c 1 -o $O/unfused python tools/quick_bench.py --nt 438 --reps 1 --spinup 2 --land-slabs 1 --no-river --no-fuse > $O/ncu_full2.log 2>&1
tail -2 $O/ncu_full2.log
