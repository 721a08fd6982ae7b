#!/bin/bash
# 8-GPU measurement suite (run under `gpurun --gpus 8 -- bash tools/n8_suite.sh 8`): bench line, phase breakdown, configs 4 / 5 (tile exchange all-reduce vs gather; instanced vs flat structure)
N=${1:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$R bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2h.json 2> gpurun_out/bench_n${N}_r2h.err
$R tools/e2e_phases.py --steps 16 --reps 5 --upload sharded --out gpurun_out/e2e_phases_n${N}_r2h.json > gpurun_out/e2e_n${N}_r2h.log 2>&1
$R tools/run_configs.py --config 4 --out gpurun_out/c4_n${N}_r2h.json > gpurun_out/c4_n${N}_r2h.log 2>&1
$R tools/run_configs.py --config 5 --out gpurun_out/c5_n${N}_r2h.json > gpurun_out/c5_n${N}_r2h.log 2>&1
$R tools/run_configs.py --config 5 --tile-collective gather --out gpurun_out/c5_n${N}_r2h_gather.json > gpurun_out/c5_n${N}_r2h_gather.log 2>&1
