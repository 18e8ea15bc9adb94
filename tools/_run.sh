python -m pytest tests -m gpu -x -q > gpurun_out/r02w_tests.log 2>&1; tail -2 gpurun_out/r02w_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02w_bench_1gpu.json 2> gpurun_out/r02w_bench_1gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/r02w_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])
for o in d['other_configs']: print(o['config'], o['ms_per_frame'], o['roofline_frac_algorithmic'])
print(d['parity'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02w_bench_reference.json 2> gpurun_out/r02w_bench_reference.err; tail -c 700 gpurun_out/r02w_bench_reference.json
