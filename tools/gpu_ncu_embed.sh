python tools/ncu_targets.py > gpurun_out/ncu_e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'embed_fwd_kernel|embed_chunk_table|embed_fwd_reduce|embed_bwd_kernel' -c 8 -o gpurun_out/prof_embed -f python tools/ncu_targets.py > gpurun_out/ncu_embed.log 2>&1
echo rc=$?
