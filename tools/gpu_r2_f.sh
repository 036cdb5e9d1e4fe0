#!/bin/bash
mkdir -p gpurun_out/r2f
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2f/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f/pytest.log
DISTBO_PROBE_SIZES=8192,4096,3000 timeout 300 python tools/probe_potrf.py > gpurun_out/r2f/probe_potrf.json 2> gpurun_out/r2f/probe_potrf.err; cat gpurun_out/r2f/probe_potrf.json
DISTBO_POTRF_RELAY2=0 DISTBO_PROBE_SIZES=8192,4096 timeout 300 python tools/probe_potrf.py > gpurun_out/r2f/probe_potrf_r1.json 2>> gpurun_out/r2f/probe_potrf.err; cat gpurun_out/r2f/probe_potrf_r1.json
timeout 120 python tools/potrf_timeline.py 4096 > gpurun_out/r2f/timeline_4096.txt 2>&1; sed -n 1,12p gpurun_out/r2f/timeline_4096.txt; sed -n 25,34p gpurun_out/r2f/timeline_4096.txt
timeout 120 python tools/potrf_timeline.py 8192 > gpurun_out/r2f/timeline_8192.txt 2>&1; sed -n 1,2p gpurun_out/r2f/timeline_8192.txt;  sed -n 40,50p gpurun_out/r2f/timeline_8192.txt
