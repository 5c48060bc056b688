#!/bin/bash
# Build-flag sweep on the GPU box: rebuild the library with different tuning macros and time it.
for cfg in "-DVIDC_MIN_BLOCKS=4" "-DVIDC_MIN_BLOCKS=5" "-DVIDC_MIN_BLOCKS=6" "-DVIDC_MIN_BLOCKS=5 -DVIDC_UNROLL=2" "-DVIDC_MIN_BLOCKS=5 -DVIDC_ROWS=8" "-DVIDC_MIN_BLOCKS=6 -DVIDC_ROWS=2" "-DVIDC_MIN_BLOCKS=3 -DVIDC_UNROLL=4"; do
  echo "=== $cfg"
  VIDC_NVCC_EXTRA="$cfg" python -m vi_depth_completion_b200.build --force > /dev/null || { echo build failed; continue; }
  python tools/quick_time.py 2>&1 | head -1 | cut -c1-330
done
python -m vi_depth_completion_b200.build --force > /dev/null
