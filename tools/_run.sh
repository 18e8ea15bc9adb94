for v in "" all8 all6; do
  if [ -n "$v" ]; then export FRAY_GPU_LIB=$PWD/fray_b200/_build/variants/libfray_gpu_$v.so; fi
  echo "== variant ${v:-main}"
  python tools/render_once.py boxed --frames 5 | tail -1
  python tools/render_once.py forest frameWidth=3840 frameHeight=2160 --frames 5 | tail -1
  python tools/render_once.py hw9/dragon --frames 5 | tail -1
done 2>&1 | tee gpurun_out/r02g_variants.log
