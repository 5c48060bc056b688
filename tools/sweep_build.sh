#!/bin/bash
# A/B of a runtime switch on the GPU box.
for v in 1 0; do
  echo "[VIDC_TILE_SKIP=$v]"
  VIDC_TILE_SKIP=$v python bench.py --steps 30 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print('  bench', round(d['value']), {n[:28]: round(v['ms'],4) for n,v in k.items()}, 'e2e', round(d['e2e']['value']))"
  VIDC_TILE_SKIP=$v python tools/roll_sweep.py 2>&1 | cut -c1-75 | sed -n '1p;3p;5p;7p'
done
