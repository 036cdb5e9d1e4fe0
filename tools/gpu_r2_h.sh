#!/bin/bash
mkdir -p gpurun_out/r2h
timeout 60 ./tools/potf2_bench | tail -1
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2h/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2h/pytest.log
DISTBO_PROBE_SIZES=8192,4096,2048,1280 timeout 300 python tools/probe_potrf.py > gpurun_out/r2h/probe_potrf.json 2> gpurun_out/r2h/probe_potrf.err; cat gpurun_out/r2h/probe_potrf.json
timeout 300 python tools/probe_small.py 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-blr > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2h/bench.json')); print({k:d[k] for k in ['value','fit_ms','fit_frac_of_fp64_peak','ms_per_step']}, d['roofline']['frac'], d['handoff']['refresh_ms'])"
