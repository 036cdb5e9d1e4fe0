#!/bin/bash
# Round-2 first GPU pass: full -m gpu suite, bench at both noise levels, full-horizon toy replay.
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -40 gpurun_out/r2a/pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --lkh-var 1e-5 --no-cpu-baseline --no-blr > gpurun_out/r2a/bench_s1e-5.json 2> gpurun_out/r2a/bench_s1e-5.err; echo "bench 1e-5 rc=$?"
timeout 1500 python tools/replay_full_horizon.py toy --impl cuda --out gpurun_out/r2a/full_horizon_toy_cuda.json > gpurun_out/r2a/replay_toy.out 2> gpurun_out/r2a/replay_toy.err; echo "replay rc=$?"
tail -3 gpurun_out/r2a/replay_toy.out
