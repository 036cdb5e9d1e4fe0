#!/bin/bash
for cfg in "6 4 -1" "6 4 28" "6 4 24" "6 4 20" "6 4 32" "4 2 28" "4 2 24" "6 4 36" "4 4 28"; do
  set -- $cfg
  DISTBO_PROBE_SIZES=8192 DISTBO_POTRF_FAR=$1 DISTBO_POTRF_FARG=$2 DISTBO_POTRF_TAIL=$3 timeout 120 python tools/probe_potrf.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['env'].get('DISTBO_POTRF_FAR'), d['env'].get('DISTBO_POTRF_FARG'), d['env'].get('DISTBO_POTRF_TAIL'), {k:(round(v['potrf_ms_min'],3), round(v['gram_plus_potrf_graph_ms'],3)) for k,v in d.items() if k!='env'}, d['8192']['rel_err'], d['8192']['info'])"
done
