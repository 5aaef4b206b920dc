#!/bin/bash
# Runs on the GPU box (gpurun): the evidence files of the final kernels -> gpurun_out/ (copied to profiles/ afterwards).
#   launch list of the bench command, ncu --set full of one forward + one gradient launch, their raw / SASS / CUDA-source pages
T=${1:-r2f}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-strong --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_forward_grid|k_gradient' -c 2 -s 2 -o gpurun_out/${T}_full -f python tools/profile_step.py 2 64 > gpurun_out/${T}_full.log 2>&1
ncu -i gpurun_out/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_full_raw.csv
ncu -i gpurun_out/${T}_full.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_full_sass.csv
ncu -i gpurun_out/${T}_full.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${T}_full_cs.csv
