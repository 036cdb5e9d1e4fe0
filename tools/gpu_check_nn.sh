timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "embed" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_regressor.py -x -q -m gpu 2>&1 | tail -8
timeout 120 python tools/probe_embed.py 2>gpurun_out/probe_embed.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(v['tensor_core']['ms']*1000,1), round(v['generic']['ms']*1000,1)) for k,v in d.items()})" || tail -5 gpurun_out/probe_embed.err
