#!/bin/bash
# A/B of library builds on the flat-batch benchmark: tools/ab_flat.sh <suffix> ... (libi2lqr<suffix>.so next to the package)
for v in "$@"; do
  I2LQR_LIB=$PWD/ilqr_iterative_tasks_b200/libi2lqr$v.so python bench.py --no-cpu --no-latency --steps 20 > gpurun_out/ab$v.json 2> gpurun_out/ab$v.err
  python - "$v" <<'P'
import json, sys
v = sys.argv[1]
d = json.loads(open("gpurun_out/ab%s.json" % v).read().strip().splitlines()[-1])
print("%-8s value %.1f M/s  frac %.4f  isolated %.4f  e2e %.1f M/s  parity %s" % (v or "(default)", d["value"] / 1e6, d["roofline"]["frac"], d["roofline"]["isolated"]["frac"], d["e2e"]["value"] / 1e6, d["e2e_matches_device_arm"]))
P
done
