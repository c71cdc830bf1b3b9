O=gpurun_out/r02u; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
for sl in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --land-slabs $sl > $O/bench_s${sl}.json 2> $O/bench_s${sl}.err; python -c "
import json; d=json.load(open('$O/bench_s${sl}.json')); k=d['roofline']['kernels']; print('slabs $sl', round(d['ms_per_step'],2), 'land', round(d['kernel_ms_per_step']['runoff_generation'],2), 'zone', round(k['zone_kernel']['ms'],2), round(k['zone_kernel']['ms_alone'],2), 'elem', round(k['element_kernel']['ms'],2), round(k['element_kernel']['ms_alone'],2), 'river', round(d['kernel_ms_per_step']['routing'],2), [round(v['ms_alone'],2) for n,v in k.items() if n.startswith('routing')], 'e2e', round(d['e2e']['ms_per_step'],2))"
done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'reach_wave' -c 3 --csv --log-file $O/wave_inst.csv python tools/quick_bench.py --nt 438 --reps 1 --spinup 2 > $O/ncu_wave.log 2>&1; grep "reach_wave" $O/wave_inst.csv | tail -4 | cut -c1-40,200-400
