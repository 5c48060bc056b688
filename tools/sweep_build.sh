#!/bin/bash
# Build-flag sweep on the GPU box: rebuild the library with different tuning macros and time it.
for ilp in 1 2; do for mb in 3 4 5; do for rows in 2 4 8; do
  cfg="-DVIDC_ILP=$ilp -DVIDC_MIN_BLOCKS=$mb -DVIDC_ROWS=$rows"
  VIDC_NVCC_EXTRA="$cfg" python -m vi_depth_completion_b200.build --force > /dev/null || { echo build failed; continue; }
  echo "$cfg $(python tools/quick_time.py 2>&1 | head -2 | python -c 'import sys,json
for l in sys.stdin:
    d=json.loads(l); print("[%s fwd %.3f inv %.3f fps %.0f]"%(d["cam"], d["forward_rgbd_mask"]["ms"], d["inverse_rot_norm"]["ms"], d["frames_per_s"]), end=" ")')"
done; done; done
python -m vi_depth_completion_b200.build --force > /dev/null
