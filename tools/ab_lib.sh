#!/bin/bash
# Same-box alternating A/B of the tree's libvpf.so against another build of the same ABI:
#   tools/ab_lib.sh <other.so> [rounds] [extra bench.py args...]
OTHER=$1; R=${2:-2}; shift 2
mkdir -p gpurun_out
for i in $(seq 1 $R); do
  for w in other tree; do
    if [ $w = other ]; then export VPF_LIB_PATH=$OTHER; else unset VPF_LIB_PATH; fi
    python bench.py --steps 8 --warmup 3 --no-extras "$@" > gpurun_out/ab_${w}_$i.json 2> gpurun_out/ab_${w}_$i.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${w}_$i.json").read())
    print("$w $i: %.2f ms/step" % d["ms_per_step"], {k: round(v["ms"], 3) for k, v in d["roofline"]["detail"].items()},
          "in-step:", {k: round(v, 2) for k, v in (d.get("encoder_breakdown_ms") or {}).items()}, flush=True)
except Exception as e:
    print("$w $i failed:", e, open("gpurun_out/ab_${w}_$i.err").read()[-1200:])
PY
  done
done
