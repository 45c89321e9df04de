#!/bin/bash
# usage: tools/ab.sh lib1.so lib2.so ...   -> kernel ms of the 10M-polyline bench for each variant (paths relative to the repo root)
for lib in "$@"; do
  TG_LIB=$PWD/$lib python bench.py --steps 10 --no-cpu-baseline $AB_FLAGS 2>&1 | python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', l['kernel_ms'], round(l['roofline']['frac'],4), l['clocks']['reasons'], l['check'])"
done
