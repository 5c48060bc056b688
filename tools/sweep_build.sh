#!/bin/bash
# Build-flag sweep on the GPU box: rebuild the library with different tuning macros and time it.
for cfg in "-DVIDC_SHEAR_MIN_FWD=20 -DVIDC_SHEAR_MIN_INV=40" "-DVIDC_SHEAR_MIN_FWD=10 -DVIDC_SHEAR_MIN_INV=30" "-DVIDC_SHEAR_MIN_FWD=35 -DVIDC_SHEAR_MIN_INV=60" "-DVIDC_SHEAR_MIN_FWD=1000 -DVIDC_SHEAR_MIN_INV=1000"; do
  VIDC_NVCC_EXTRA="$cfg" python -m vi_depth_completion_b200.build --force > /dev/null || { echo build failed; continue; }
  echo "[$cfg] $(python tools/quick_time.py 2>/dev/null | head -3 | python -c 'import sys,json
for l in sys.stdin:
    d=json.loads(l); print("[%s fwd %.3f inv %.3f fps %.0f]"%(d["cam"], d["forward_rgbd_mask"]["ms"], d["inverse_rot_norm"]["ms"], d["frames_per_s"]), end=" ")')"
  python tools/roll_sweep.py 2>&1 | cut -c1-75
done
python -m vi_depth_completion_b200.build --force > /dev/null
