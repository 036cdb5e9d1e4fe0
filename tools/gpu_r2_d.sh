#!/bin/bash
mkdir -p gpurun_out/r2d
echo "--- potf2_bench (2x2 chain + in-block look-ahead)"; timeout 60 ./tools/potf2_bench | tail -2
echo "--- potf2_bench_1x1 (round-1 flow)"; timeout 60 ./tools/potf2_bench_1x1 | tail -2
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d/pytest.log
DISTBO_PROBE_SIZES=8192,4096,3000,2048,1280 timeout 300 python tools/probe_potrf.py > gpurun_out/r2d/probe_potrf.json 2> gpurun_out/r2d/probe_potrf.err; cat gpurun_out/r2d/probe_potrf.json
DISTBO_POTRF_SMALL=0 DISTBO_PROBE_SIZES=2048,1280 timeout 300 python tools/probe_potrf.py > gpurun_out/r2d/probe_potrf_nosmall.json 2>> gpurun_out/r2d/probe_potrf.err; cat gpurun_out/r2d/probe_potrf_nosmall.json
timeout 300 python tools/probe_small.py > gpurun_out/r2d/probe_small.txt 2>&1; cat gpurun_out/r2d/probe_small.txt | tail -5
