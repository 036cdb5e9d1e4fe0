#!/bin/bash
mkdir -p gpurun_out/r2b
./tools/factor32_bench > gpurun_out/r2b/factor32.txt 2>&1; cat gpurun_out/r2b/factor32.txt
./tools/potf2_bench > gpurun_out/r2b/potf2.txt 2>&1; tail -2 gpurun_out/r2b/potf2.txt
python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
tail -30 gpurun_out/r2b/pytest.log
timeout 900 python tools/replay_full_horizon.py toy --impl cuda --teacher-forced --out gpurun_out/r2b/full_horizon_toy_cuda_teacher_forced.json > gpurun_out/r2b/replay_tf.out 2> gpurun_out/r2b/replay_tf.err; echo "replay rc=$?"
tail -2 gpurun_out/r2b/replay_tf.out
timeout 900 python tools/replay_full_horizon.py toy --impl cuda --out gpurun_out/r2b/full_horizon_toy_cuda.json > gpurun_out/r2b/replay.out 2> gpurun_out/r2b/replay.err; echo "replay rc=$?"
tail -2 gpurun_out/r2b/replay.out
