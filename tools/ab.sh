#!/bin/bash
# usage: tools/ab.sh <variant>...   runs a short bench with each variants/libs2m_<v>.so, prints value / e2e / K4 us
for v in "$@"; do
  S2M_LIB=$PWD/sc-a-loam_b200/csrc/variants/libs2m_$v.so timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-sharded --no-os1 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    j = json.loads(open("gpurun_out/ab_%s.json" % v).read().strip().splitlines()[-1])
    print(v, "value %.0f e2e %.0f k4_us %.1f frac %.4f" % (j["value"], j["e2e"]["value"], j["roofline"]["avg_launch_us"], j["roofline"]["frac"]), j["config"]["phase_ms_per_step_single_lane"])
except Exception as e:
    print(v, "FAILED", e)
PY
done
