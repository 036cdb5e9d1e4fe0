for cfg in "2 4" "2 2" "3 4" "4 4" "2 8" "3 8" "2 6"; do
  set -- $cfg
  DISTBO_POTRF_FAR=$1 DISTBO_POTRF_FARG=$2 timeout 120 python tools/probe_potrf.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['env'], {k:round(v['potrf_ms_min'],3) for k,v in d.items() if k!='env'}, d['8192']['rel_err'], d['8192']['info'])"
done
