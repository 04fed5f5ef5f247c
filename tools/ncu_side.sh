#!/bin/bash
# launch lists (ncu --metrics gpu__time_duration.sum, cold caches, serialised) of the optimiser kernels: LocalBA windows (cluster kernel of the
# batched mode, whole-GPU kernel of a single window) and one pose call
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_lba_fused -c 8 --csv --log-file gpurun_out/r2ao_lba_fused.csv python tools/lba_time.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_lba_grid -c 8 --csv --log-file gpurun_out/r2ao_lba_grid.csv python tools/lba_time.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/r2ao_pose_launches.csv python tools/pose_one.py > /dev/null 2>&1
true
