FRAY_DIST_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-other-configs > gpurun_out/r02h_bench_8gpu.json 2> gpurun_out/r02h_bench_8gpu.err; grep -v '^\*\|OMP_NUM' gpurun_out/r02h_bench_8gpu.err | tail -12; cat gpurun_out/r02h_bench_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('value','ms_per_step','n_gpus','e2e'): print(k, d.get(k))"
