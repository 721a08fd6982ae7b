#!/bin/bash
# ncu evidence of round 2 (one GPU, under gpurun): launch list of the bench command + one `--set full` capture of the two hot kernels.
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 2 --warmup 3 --no-torch-baseline --no-cpu-baseline --no-strong --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2a.csv $CMD > gpurun_out/launches_r2a.log 2>&1
# bounces 0..3 of one warmed-up step: skip the first 5 steps' launches of each kernel (4 per step)
ncu --set full --clock-control none --import-source on -k regex:k_extend_cw -s 20 -c 4 -o gpurun_out/prof_extend_cw_r2a -f $CMD > gpurun_out/prof_extend_r2a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 20 -c 4 -o gpurun_out/prof_shade_r2a -f $CMD > gpurun_out/prof_shade_r2a.log 2>&1
