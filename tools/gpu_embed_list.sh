timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/embed_launches.csv python tools/probe_embed.py > gpurun_out/embed_list.log 2>&1
echo rc=$?
