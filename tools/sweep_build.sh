#!/bin/bash
# Build-flag sweep on the GPU box: rebuild the library with different tuning macros and time it.
for rows in 2 4; do
  cfg="-DVIDC_TMA_ROWS=$rows"
  VIDC_NVCC_EXTRA="$cfg" python -m vi_depth_completion_b200.build --force > /dev/null || { echo build failed; continue; }
  echo "$cfg $(VIDC_TMA=1 python tools/quick_time.py 2>&1 | head -2 | python -c 'import sys,json
for l in sys.stdin:
    d=json.loads(l); print("[%s fwd %.3f inv %.3f fps %.0f]"%(d["cam"], d["forward_rgbd_mask"]["ms"], d["inverse_rot_norm"]["ms"], d["frames_per_s"]), end=" ")')"
done
python -m vi_depth_completion_b200.build --force > /dev/null
