#!/bin/bash
mkdir -p gpurun_out/r2c
python -m pytest tests/test_gpu_round2.py -m gpu -q -s > gpurun_out/r2c/pytest_round2.log 2>&1; echo "pytest rc=$?"
grep -E "polish |ill-conditioned|passed|failed" gpurun_out/r2c/pytest_round2.log
for c in config1 config4; do
  timeout 900 python bench.py --workload $c > gpurun_out/r2c/bench_$c.json 2> gpurun_out/r2c/bench_$c.err; echo "bench $c rc=$?"
  tail -c 1500 gpurun_out/r2c/bench_$c.json; tail -3 gpurun_out/r2c/bench_$c.err
done
