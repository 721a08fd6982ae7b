#!/bin/bash
# Final check of the round on one B200 (under gpurun): GPU test suite, smoke(), the default bench line, configs 4 / 5 at full size on one GPU and
# the ncu launch list of the bench command.  Outputs under gpurun_out/final5_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/final5
timeout 300 python -m pytest tests -m gpu -x -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py > ${P}_bench_n1.json 2> ${P}_bench_n1.err; echo "bench rc=$?"
timeout 150 python tools/run_configs.py --config 4 --out ${P}_c4_n1.json > /dev/null 2> ${P}_c4.err; echo "c4 rc=$?"
timeout 200 python tools/run_configs.py --config 5 --out ${P}_c5_n1.json > /dev/null 2> ${P}_c5.err; echo "c5 rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file ${P}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-torch-baseline --no-cpu-baseline --no-strong --e2e-steps 1 > ${P}_launches.log 2>&1; echo "ncu rc=$?"
tail -3 ${P}_pytest.log; cat ${P}_smoke.log | tail -2
