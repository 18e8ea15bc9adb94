set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.log 2>&1; tail -3 gpurun_out/r02b_tests.log
python tools/time_configs.py --fp32 2>&1 | tee gpurun_out/r02b_configs.log
for sc in boxed hw9/dragon; do n=$(echo $sc | tr / _); ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02b_launches_$n.csv python tools/render_once.py $sc --frames 2 > /dev/null 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02b_launches_forest.csv python tools/render_once.py forest frameWidth=3840 frameHeight=2160 --frames 2 > /dev/null 2>&1
