#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_su|k_cells_mid|k_cells_slow|k_cells_coh|k_cells_fast" --launch-skip 24 -c 5 -o gpurun_out/prof_r02 -f python tools/profile_target.py > gpurun_out/prof_r02.log 2>&1
tail -3 gpurun_out/prof_r02.log
ls -la gpurun_out/prof_r02.ncu-rep
