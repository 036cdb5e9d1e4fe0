#!/bin/bash
# ncu --set full capture of the BLR kernels on the config-5 shaped probe (tools/probe_blr.py), after a plain run.
set -u
mkdir -p gpurun_out
python tools/probe_blr.py > gpurun_out/probe_blr.json 2> gpurun_out/probe_blr.err; echo "plain rc=$?"
cat gpurun_out/probe_blr.json
ncu --set full --clock-control none --import-source on -k regex:blr_ -s 30 -c 24 -o gpurun_out/prof_blr -f \
  python tools/probe_blr.py > gpurun_out/ncu_blr.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_blr.log
