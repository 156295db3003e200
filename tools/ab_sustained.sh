#!/bin/bash
# DEVELOPMENT: sustained (power-capped) bench line of several builds of the library: tools/ab_sustained.sh default w_scalar ...
for spec in "$@"; do
  l="${spec%%:*}"; envs=""; [ "$spec" != "$l" ] && envs="${spec#*:}"
  L=""; [ "$l" != default ] && L=/root/repo/tools/_build/libnpde_$l.so
  env $envs NPDE_B200_LIB=$L python bench.py --no-cpu --no-ttr --no-train > gpurun_out/sus_$l.json 2>/dev/null
  python - "$l" "$spec" <<'PY'
import json,sys
l=sys.argv[1]
d=json.load(open("gpurun_out/sus_%s.json"%l))
print("%-28s value %.1f G  frac %.4f  launch %.2f us  sm %s MHz  e2e %.1f G"%(sys.argv[2],d["value"]/1e9,d["roofline"]["frac"],d["roofline"]["avg_launch_ms"]*1e3,d["clocks"]["sm_mhz"],d["e2e"]["value"]/1e9))
PY
done
