O=gpurun_out/r03u; mkdir -p $O
python bench.py --config 2 --steps 10 --warmup 3 > $O/bench_config2_n1.json 2> $O/bench_config2_n1.err; echo rc=$?; tail -2 $O/bench_config2_n1.err; python -c "
import json; d=json.load(open('$O/bench_config2_n1.json')); k=d['roofline']['kernels']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('cpu_baseline'), d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['frac_alone'], d['config']['recstep_executed_frac'], k['element_kernel']['ms_alone'])"
