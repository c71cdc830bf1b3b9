O=gpurun_out/r03a; mkdir -p $O
for v in "" _eb16; do for f in "" "--no-land-fuse"; do for wb in 3 4; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline --wave-blocks $wb $f > $O/bench$v$f$wb.json 2> $O/bench$v$f$wb.err; python -c "
import json; d=json.load(open('$O/bench$v$f$wb.json')); k=d['roofline']['kernels']; print('lib$v $f wb$wb', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'zone', round(k['zone_kernel']['ms'],2), 'elem', round(k['element_kernel']['ms'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done; done; done
