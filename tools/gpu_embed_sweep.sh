timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "embed" 2>&1 | tail -3
DISTBO_EMBED_TC_MODE=mt2 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "embed_forward_tensor" 2>&1 | tail -3
DISTBO_EMBED_TC_MODE=mt1 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "embed_forward_tensor" 2>&1 | tail -3
for cfg in "pair 7 0" "pair 6 0" "pair 8 0" "mt2 7 0" "pair 7 1"; do
  set -- $cfg
  echo "MODE=$1 teams=$2 skip=$3"; DISTBO_EMBED_TC_SKIP=$3 DISTBO_EMBED_TC_MODE=$1 DISTBO_EMBED_TC_WARPS=$2 timeout 120 python tools/probe_embed.py 2>gpurun_out/probe_embed.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(v['tensor_core']['ms']*1000,1), round(v['generic']['ms']*1000,1)) for k,v in d.items()})" || tail -5 gpurun_out/probe_embed.err
done
