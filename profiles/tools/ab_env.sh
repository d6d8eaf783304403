#!/bin/bash
# In-box A/B of one environment switch of libewk (GPU instances of the pool differ by up to ~10 % on K3, so only runs on
# the same box compare):  gpurun -- 'bash profiles/tools/ab_env.sh VAR valueA valueB [rounds] [steps]'
VAR=$1; A=$2; B=$3; R=${4:-2}; S=${5:-20}
for i in $(seq 1 $R); do
  for v in $A $B; do
    env $VAR=$v python bench.py --no-cpu --steps $S > gpurun_out/ab_${VAR}_${v}_$i.json 2>/dev/null
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_${VAR}_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items()},
          "e2e", round(d["e2e"]["value"] / 1e6, 3))
PY
