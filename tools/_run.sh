for v in "" z3 z5 z6; do
  if [ -n "$v" ]; then export FRAY_GPU_LIB=$PWD/fray_b200/_build/variants/libfray_gpu_$v.so; fi
  echo "== variant ${v:-main(4)}"
  python tools/render_once.py zaphod --frames 6 | sort -k6 -n | head -1
done 2>&1 | tee gpurun_out/r02v_zaphod.log
