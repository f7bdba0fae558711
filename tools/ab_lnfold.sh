#!/bin/bash
# Same-box A/B of the LayerNorm folding modes (env VPF_LN_FOLD): alternating bench.py runs.
#   tools/ab_lnfold.sh [rounds] [extra bench.py args...]
R=${1:-2}; shift
mkdir -p gpurun_out
for i in $(seq 1 $R); do
  for which in ${VARIANTS:-0 kernel epilogue}; do
    VPF_LN_FOLD=$which python bench.py --steps 8 --warmup 3 --no-extras "$@" > gpurun_out/lnfold_${which}_$i.json 2> gpurun_out/lnfold_${which}_$i.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/lnfold_${which}_$i.json").read())
    r = d["roofline"]["detail"]
    print("ln_fold=$which $i: %.2f ms/step  gemm frac %.3f" % (d["ms_per_step"], d["roofline"]["frac"]),
          {k: round(v["ms"], 3) for k, v in r.items()},
          "in-step:", {k: round(v, 2) for k, v in (d.get("encoder_breakdown_ms") or {}).items()}, flush=True)
except Exception as e:
    print("ln_fold=$which $i failed:", e, open("gpurun_out/lnfold_${which}_$i.err").read()[-1500:])
PY
  done
done
