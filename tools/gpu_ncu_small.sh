python tools/probe_small.py > gpurun_out/small_plain.log 2>&1 &&
DISTBO_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/small_launches.csv python tools/probe_small.py > gpurun_out/small_ncu.log 2>&1
echo rc=$?
