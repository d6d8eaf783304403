#!/bin/bash
# In-box A/B of two builds of libewk.so (GPU instances of the pool differ by up to ~10 % on K3, so only runs on the same
# box compare): alternates `bench.py --no-cpu` between the in-tree library and experiments/libewk_alt.so.
#   gpurun -- 'bash profiles/tools/ab_lib.sh [rounds] [steps]'
R=${1:-2}; S=${2:-20}
cp easywakeword_b200/libewk.so /tmp/libewk_new.so
for i in $(seq 1 $R); do
  for v in new alt; do
    if [ $v = new ]; then cp /tmp/libewk_new.so easywakeword_b200/libewk.so; else cp experiments/libewk_alt.so easywakeword_b200/libewk.so; fi
    python bench.py --no-cpu --steps $S > gpurun_out/ab_${v}_$i.json 2>/dev/null
  done
done
cp /tmp/libewk_new.so easywakeword_b200/libewk.so
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items()},
          "dense", round(d["dense"]["kernel_ms"], 3))
PY
