for c in "" 2 4 8; do
  if [ -n "$c" ]; then export FRAY_GPU_CHUNK=$c; fi
  echo "== chunk ${c:-auto}"
  python tools/render_once.py zaphod --frames 6 | sort -k6 -n | head -1
  python tools/render_once.py cornell_box pathsPerPixel=40 --frames 6 | sort -k6 -n | head -1
done
