#!/bin/bash
# One gpurun call: GPU parity tests, the headline bench, then ncu (launch list + full capture of the
# top kernels) on a reduced-candidate run of the same bench command.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" = "0" ]; then
SMALL="python bench.py --candidates 75776 --steps 1 --warmup 1 --no-cpu-baseline --no-blr --fit-epochs 1"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$SMALL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 4 -c 2 -o gpurun_out/prof_score -f $SMALL > gpurun_out/ncu_score.log 2>&1
echo "ncu score rc=$?"
fi
