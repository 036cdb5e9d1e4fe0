#!/bin/bash
mkdir -p gpurun_out/r2n2
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "second_device or two_gpu" > gpurun_out/r2n2/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2n2/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2n2/bench_n2.json 2> gpurun_out/r2n2/bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2n2/bench_n2.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ['value','n_gpus','fit_ms','ms_per_step','handoff','best']}, d['e2e']['value'])"
tail -3 gpurun_out/r2n2/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-sample 4096 > gpurun_out/r2n2/bench_ref_n2.json 2> gpurun_out/r2n2/bench_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-300 gpurun_out/r2n2/bench_ref_n2.json
