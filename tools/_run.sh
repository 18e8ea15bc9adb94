python -m pytest tests -m gpu -x -q > gpurun_out/r02n_tests.log 2>&1; tail -3 gpurun_out/r02n_tests.log
for v in "" 1; do
  if [ -n "$v" ]; then export FRAY_GPU_NO_FUSE=1; fi
  echo "== nofuse=${v:-0}"
  python tools/render_once.py forest frameWidth=3840 frameHeight=2160 --frames 5 | tail -1
  python tools/render_once.py forest frameWidth=3840 frameHeight=2160 wantAA=on --frames 4 | tail -1
  python tools/render_once.py hw9/dragon --frames 5 | tail -1
  python tools/render_once.py forest frameWidth=1920 frameHeight=1080 --frames 5 | tail -1
done 2>&1 | tee gpurun_out/r02n_fused.log
