#!/bin/bash
i=0
for v in "$@"; do
  env $v python bench_configs.py episodes 4096 > gpurun_out/epab$i.json 2> gpurun_out/epab$i.err
  python - "$v" $i <<'P'
import json, sys
v, i = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open("gpurun_out/epab%s.json" % i).read().strip().splitlines()[-1])
    print("%-36s" % v, " ".join("%.3f" % r["wall_s"] for r in d["laps"]), "setup %.3f" % d["setup_s"])
except Exception as ex:
    print(v, "FAILED", ex, open("gpurun_out/epab%s.err" % i).read()[-300:])
P
  i=$((i+1))
done
