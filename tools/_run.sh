python -m pytest tests -m gpu -x -q > gpurun_out/r02y_tests.log 2>&1; tail -2 gpurun_out/r02y_tests.log
python tools/render_once.py boxed --frames 6 | sort -k6 -n | head -1
python tools/render_once.py forest frameWidth=3840 frameHeight=2160 --frames 5 | sort -k6 -n | head -1
python tools/render_once.py hw9/dragon --frames 5 | sort -k6 -n | head -1
