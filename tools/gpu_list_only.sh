SMALL="python bench.py --candidates 75776 --steps 1 --warmup 1 --no-cpu-baseline --no-blr --fit-epochs 1"
$SMALL > gpurun_out/plain.log 2>&1 && tail -c 600 gpurun_out/plain.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; grep -c score_kernel gpurun_out/launches.csv
