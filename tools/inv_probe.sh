#!/bin/bash
# Development helper: timing of the fused calls + instruction count / duration of the inverse kernel under ncu (two metrics, one pass)
python tools/quick_time.py one
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"unwarp_normals|frame_params_inv" -c 2 python tools/profile_once.py S2 256 1 2>&1 | grep -E "unwarp_normals|frame_params|inst_executed|time_duration|issue_active|wavefronts|bank_conflicts|warps_active"
