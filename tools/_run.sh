for sp in samples auto tiles; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --split $sp --no-other-configs > gpurun_out/r02m_bench_4gpu_$sp.json 2> gpurun_out/r02m_bench_4gpu_$sp.err; cat gpurun_out/r02m_bench_4gpu_$sp.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$sp', 'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'], d['config']['split'], d.get('multi_gpu_check',{}).get('max_abs_diff'))"
done
