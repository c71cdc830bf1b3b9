O=gpurun_out/r02z; mkdir -p $O
for v in "" _eb14 _eb16; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python tools/quick_bench.py --nt 438 --reps 3 --spinup 1 --land-slabs 1 --no-river > $O/quick$v.log 2>&1; echo "lib$v"; grep "^rep [2]" $O/quick$v.log | cut -c1-100
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wave-blocks 3 > $O/bench$v.json 2> $O/bench$v.err; python -c "
import json; d=json.load(open('$O/bench$v.json')); k=d['roofline']['kernels']; print('lib$v', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'zone', round(k['zone_kernel']['ms'],2), round(k['zone_kernel']['ms_alone'],2), 'elem', round(k['element_kernel']['ms'],2), round(k['element_kernel']['ms_alone'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), [round(v['ms_alone'],2) for n,v in k.items() if n.startswith('routing')], 'e2e', round(d['e2e']['ms_per_step'],2))"
done
