set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench_1gpu.json 2> gpurun_out/r02k_bench_1gpu.err; tail -c 600 gpurun_out/r02k_bench_1gpu.json
python tools/time_configs.py --fp32 > gpurun_out/r02k_all_configs_timing.log 2>&1
python tools/share_time.py --world 8 --chunks 0,1,2 > gpurun_out/r02k_share_times.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:renderKernel -s 3 -c 1 -f -o gpurun_out/r02k_cornell256 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_x.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02k_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_x.log 2>&1
for k in waveTraceKernel waveShadeKernel waveShadowKernel; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02k_forest_$k python tools/render_once.py forest frameWidth=3840 frameHeight=2160 --frames 2 > gpurun_out/ncu_x.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:waveShadowKernel -s 1 -c 1 -f -o gpurun_out/r02k_boxed_waveShadowKernel python tools/render_once.py boxed --frames 2 > gpurun_out/ncu_x.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:waveTraceKernel -s 6 -c 1 -f -o gpurun_out/r02k_dragon_waveTraceKernel python tools/render_once.py hw9/dragon --frames 2 > gpurun_out/ncu_x.log 2>&1
for sc in boxed hw9/dragon; do n=$(echo $sc | tr / _); ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02k_launches_$n.csv python tools/render_once.py $sc --frames 2 > /dev/null 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02k_launches_forest4k.csv python tools/render_once.py forest frameWidth=3840 frameHeight=2160 --frames 2 > /dev/null 2>&1
ls -la gpurun_out | tail -20
