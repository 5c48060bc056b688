#!/bin/bash
# Build-flag sweep on the GPU box: rebuild the library with different tuning macros and time it.
for mb in 3 4 5 6; do for un in 1 2 4; do
  cfg="-DVIDC_MIN_BLOCKS=$mb -DVIDC_UNROLL=$un"
  VIDC_NVCC_EXTRA="$cfg" python -m vi_depth_completion_b200.build --force > /dev/null || { echo build failed; continue; }
  echo "$cfg $(python tools/quick_time.py one 2>&1 | head -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("fwd %.3f inv %.3f fps %.0f"%(d["forward_rgbd_mask"]["ms"], d["inverse_rot_norm"]["ms"], d["frames_per_s"]))')"
done; done
python -m vi_depth_completion_b200.build --force > /dev/null
