python -m pytest tests -m gpu -x -q > gpurun_out/r02t_tests.log 2>&1; tail -2 gpurun_out/r02t_tests.log
python tools/time_configs.py --fp32 2>&1 | tee gpurun_out/r02t_all_configs_timing.log
ncu --set full --clock-control none --import-source on -k regex:waveShadowKernel -s 1 -c 1 -f -o gpurun_out/r02t_boxed_waveShadowKernel python tools/render_once.py boxed --frames 2 > gpurun_out/ncu_x.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02t_launches_boxed.csv python tools/render_once.py boxed --frames 2 > /dev/null 2>&1
