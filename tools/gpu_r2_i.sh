#!/bin/bash
mkdir -p gpurun_out/r2i
for c in config2 config3; do
  timeout 1500 python bench.py --workload $c > gpurun_out/r2i/bench_$c.json 2> gpurun_out/r2i/bench_$c.err; echo "bench $c rc=$?"
  tail -c 2500 gpurun_out/r2i/bench_$c.json; tail -3 gpurun_out/r2i/bench_$c.err
done
