python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02p_bench_8gpu.json 2> gpurun_out/r02p_bench_8gpu.err; cat gpurun_out/r02p_bench_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'], d['config']['split'], d['config'].get('e2e_split'), d.get('multi_gpu_check',{}).get('max_abs_diff'))
sp=d.get('single_process') or {}
for k in ('tiles','samples'): print(k, sp.get(k,{}).get('e2e_ms_per_frame'), sp.get(k,{}).get('slowest_share_kernel_ms'), sp.get(k,{}).get('bit_identical_to_one_gpu'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02p_bench_2gpu.json 2> gpurun_out/r02p_bench_2gpu.err; cat gpurun_out/r02p_bench_2gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'], d['config']['split'])"
