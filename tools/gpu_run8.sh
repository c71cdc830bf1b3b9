O=gpurun_out/r02i; mkdir -p $O
for rep in 1 2; do for v in "" _nobackoff _backoff6; do
HYDPY_B200_LIB=$PWD/hydpy_b200/libhydpy_b200$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench$v.$rep.json 2> $O/bench$v.$rep.err
python -c "
import json; d=json.load(open('$O/bench$v.$rep.json')); k=d['roofline']['kernels']; print('lib$v rep $rep', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), 'river alone', round(k['routing (reach_head + reach_bulk + reach_wave kernels)']['ms_alone'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done; done
