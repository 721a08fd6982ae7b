#!/bin/bash
# second 8-GPU pass of round 2: bench line (strong-scaling record with result_rank + coalesced upload), phase breakdown, config 5 with instancing
N=${1:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$R bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_r2c.json 2> gpurun_out/bench_n${N}_r2c.err
$R tools/e2e_phases.py --steps 16 --upload sharded --out gpurun_out/e2e_phases_n${N}_r2c.json > gpurun_out/e2e_n${N}_r2c.log 2>&1
$R tools/e2e_phases.py --steps 16 --upload sharded --all-ranks-read --out gpurun_out/e2e_phases_n${N}_r2c_allread.json > gpurun_out/e2e_n${N}_r2c_allread.log 2>&1
$R bench.py --gpus $N --steps 20 --warmup 5 --no-strong > gpurun_out/bench_n${N}_r2c_b.json 2> gpurun_out/bench_n${N}_r2c_b.err
