#!/bin/bash
# Same-box A/B of two builds of libvpf.so (same ABI): alternating bench.py runs, previous build
# (ab/libvpf_prev.so, built from the last commit by the developer) against the tree's.
#   tools/ab_bench.sh [rounds] [extra bench.py args...]
R=${1:-2}; shift
mkdir -p gpurun_out
for i in $(seq 1 $R); do
  for which in prev new; do
    if [ $which = prev ]; then export VPF_LIB_PATH=$PWD/ab/libvpf_prev.so; else unset VPF_LIB_PATH; fi
    python bench.py --steps 8 --warmup 3 --no-extras "$@" > gpurun_out/ab_${which}_$i.json 2> gpurun_out/ab_${which}_$i.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_${which}_$i.json").read())
r = d["roofline"]["detail"]
print("$which $i: %.2f ms/step  gemm frac %.3f  crop %.3f ms" % (d["ms_per_step"], d["roofline"]["frac"], d["hbm_kernels"]["crop_resize_patchify"]["ms"]),
      {k: round(v["ms"], 3) for k, v in r.items()},
      "in-step:", {k: round(v, 2) for k, v in (d.get("encoder_breakdown_ms") or {}).items()})
PY
  done
done
