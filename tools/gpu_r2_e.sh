#!/bin/bash
mkdir -p gpurun_out/r2e
timeout 120 python tools/potrf_timeline.py 4096 > gpurun_out/r2e/timeline_4096.txt 2>&1; head -40 gpurun_out/r2e/timeline_4096.txt
timeout 120 python tools/potrf_timeline.py 8192 > gpurun_out/r2e/timeline_8192.txt 2>&1; sed -n 1,3p gpurun_out/r2e/timeline_8192.txt; sed -n 30,70p gpurun_out/r2e/timeline_8192.txt
