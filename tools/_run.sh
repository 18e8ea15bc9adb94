for v in "" g5 g6 g8; do
  if [ -n "$v" ]; then export FRAY_GPU_LIB=$PWD/fray_b200/_build/variants/libfray_gpu_$v.so; fi
  echo "== variant ${v:-main(7)}"
  python tools/render_once.py cornell_box pathsPerPixel=256 --frames 5 | sort -k6 -n | head -1
  python tools/render_once.py smallpt pathsPerPixel=256 --frames 4 | sort -k6 -n | head -1
done 2>&1 | tee gpurun_out/r02x_gi.log
