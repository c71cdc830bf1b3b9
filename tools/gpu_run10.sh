O=gpurun_out/r02l; mkdir -p $O
for v in _base ""; do for sl in 1 2; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --land-slabs $sl > $O/bench${v}_s$sl.json 2> $O/bench${v}_s$sl.err; python -c "
import json; d=json.load(open('$O/bench${v}_s$sl.json')); k=d['roofline']['kernels']; print('lib$v slabs $sl', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'zone', round(k['zone_kernel']['ms'],2), round(k['zone_kernel']['ms_alone'],2), 'elem', round(k['element_kernel']['ms'],2), round(k['element_kernel']['ms_alone'],2), 'river', round(d['kernel_ms_per_step']['routing'],2))"
done; done
for v in _base ""; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python tools/quick_bench.py --nt 438 --reps 3 --spinup 0 > $O/quick438$v.log 2>&1; grep "^rep [012]" $O/quick438$v.log | cut -c1-250
done
