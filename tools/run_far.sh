timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "potrf or cholesky" 2>&1 | tail -3
for cfg in "4 4 0" "6 4 0" "8 4 0" "12 4 0" "6 4 1" "8 4 1" "8 2 0"; do
  set -- $cfg
  DISTBO_POTRF_FAR=$1 DISTBO_POTRF_FARG=$2 DISTBO_POTRF_CHAIN=$3 timeout 120 python tools/probe_potrf.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['env'], {k:(round(v['potrf_ms_min'],3), round(v['gram_plus_potrf_graph_ms'],3)) for k,v in d.items() if k!='env'}, d['8192']['rel_err'], d['8192']['info'])"
done
