timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'embed_fwd_tc_kernel' -c 1 -o gpurun_out/prof_embed_tc -f python tools/probe_embed.py > gpurun_out/ncu_embed_tc.log 2>&1
echo rc=$?
tail -3 gpurun_out/ncu_embed_tc.log
