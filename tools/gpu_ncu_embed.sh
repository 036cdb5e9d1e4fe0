python tools/ncu_targets.py > gpurun_out/ncu_e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'embed_fwd_kernel<float' -c 1 -o gpurun_out/prof_embed32 -f python tools/ncu_targets.py > gpurun_out/ncu_embed32.log 2>&1
echo rc=$?
