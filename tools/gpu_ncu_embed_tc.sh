timeout 120 python tools/probe_embed.py > gpurun_out/probe_embed.json 2> gpurun_out/probe_embed.err; echo "probe rc=$?"; cat gpurun_out/probe_embed.json
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'embed_fwd_tc_kernel' -c 1 -o gpurun_out/prof_embed_tc -f python tools/probe_embed.py > gpurun_out/ncu_embed_tc.log 2>&1
echo rc=$?
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/embed_launches.csv python tools/probe_embed.py > gpurun_out/embed_list.log 2>&1
grep -E "embed_" gpurun_out/embed_launches.csv | head -3 | cut -d, -f5,15-
