timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "embed" > gpurun_out/pytest_embed_tc.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_embed_tc.log
timeout 120 python tools/probe_embed.py > gpurun_out/probe_embed.json 2> gpurun_out/probe_embed.err; echo "probe rc=$?"
cat gpurun_out/probe_embed.json; tail -5 gpurun_out/probe_embed.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/embed_launches.csv python tools/probe_embed.py > gpurun_out/embed_list.log 2>&1
grep -E "embed_" gpurun_out/embed_launches.csv | head -6 | cut -d, -f5,15-
