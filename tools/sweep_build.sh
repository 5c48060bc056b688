#!/bin/bash
# Build-flag sweep on the GPU box: rebuild the library with different tuning macros and time it.
for cfg in "-DVIDC_SHEAR_BLOCKS_FWD=6" "-DVIDC_SHEAR_BLOCKS_FWD=5"; do
  VIDC_NVCC_EXTRA="$cfg" python -m vi_depth_completion_b200.build --force > /dev/null || { echo build failed; continue; }
  echo "[$cfg]"
  python bench.py --steps 30 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print('  bench', round(d['value']), {n[:28]: round(v['ms'],4) for n,v in k.items()})"
  python tools/roll_sweep.py 2>&1 | cut -c1-75
done
python -m vi_depth_completion_b200.build --force > /dev/null
