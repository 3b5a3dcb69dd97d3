#!/bin/bash
# A/B of environment switches on the flat-batch benchmark: tools/ab_env.sh "I2LQR_LEAN=0" "I2LQR_LEAN=1" ...
i=0
for v in "$@"; do
  env $v python bench.py --no-cpu --no-latency --steps 20 > gpurun_out/abenv$i.json 2> gpurun_out/abenv$i.err
  python - "$v" $i <<'P'
import json, sys
v, i = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open("gpurun_out/abenv%s.json" % i).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-28s value %.1f M/s  frac %.4f  isolated %.4f (%.3f ms)  e2e %.1f M/s  parity %s" % (v, d["value"] / 1e6, r["frac"], r["isolated"]["frac"], r["isolated"]["launch_ms"], d["e2e"]["value"] / 1e6, d["e2e_matches_device_arm"]))
except Exception as ex:
    print(v, "FAILED", ex, open("gpurun_out/abenv%s.err" % i).read()[-400:])
P
  i=$((i+1))
done
