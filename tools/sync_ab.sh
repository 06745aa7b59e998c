#!/bin/bash
# usage: tools/sync_ab.sh  -- host-wait mode A/B: all cores vs 4 cores (what a rank gets on a 32-core 8-GPU host)
nproc
for cfg in "all spin" "all block" "4 spin" "4 block" "4 auto"; do
  set -- $cfg
  pre=""; [ "$1" = "4" ] && pre="taskset -c 0-3"
  S2M_SYNC=$2 $pre timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-sharded --no-os1 > gpurun_out/sync_$1_$2.json 2> gpurun_out/sync_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.loads(open("gpurun_out/sync_%s_%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
    print(sys.argv[1], sys.argv[2], "value %.0f e2e %.0f host_wall %.2f ms_per_step %.2f" % (j["value"], j["e2e"]["value"], j["config"]["host_wall_ms_per_step"], j["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
done
