#!/bin/bash
# The 8-GPU measurement suite of round 2 (run under `gpurun --gpus 8`): bench line, per-phase e2e breakdown (sharded vs direct scene upload),
# configs 4 and 5 at full size (tile exchange: gather vs all-reduce).  Results land in gpurun_out/.
N=${1:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$R bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
$R tools/e2e_phases.py --steps 16 --upload sharded --out gpurun_out/e2e_phases_n${N}_sharded.json > gpurun_out/e2e_n${N}s.log 2>&1
$R tools/e2e_phases.py --steps 16 --upload direct --out gpurun_out/e2e_phases_n${N}_direct.json > gpurun_out/e2e_n${N}d.log 2>&1
$R tools/run_configs.py --config 4 --out gpurun_out/c4_n$N.json > gpurun_out/c4_n$N.log 2>&1
$R tools/run_configs.py --config 5 --out gpurun_out/c5_n$N.json > gpurun_out/c5_n$N.log 2>&1
$R tools/run_configs.py --config 5 --tile-collective allreduce --out gpurun_out/c5_n${N}_allreduce.json > gpurun_out/c5_n${N}_ar.log 2>&1
nvidia-smi topo -m > gpurun_out/topo$N.txt 2>&1
