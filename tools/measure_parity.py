"""Records the achieved tracker-level parity figures (tests/parity_cases.py) for every case,
operand type and residual-stream type -> gpurun_out/parity.json (copied to profiles/).  The
tolerances the tests assert (TOL) are chosen from this table with head-room.

    python tools/measure_parity.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_cases import CASES, TOL, gpu_metrics  # noqa: E402


def main():
    out = {"tolerances_asserted": TOL, "rows": []}
    for case in CASES:
        for op, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
            for res, rdt in (("f32", torch.float32), ("f16", torch.float16)):
                m = gpu_metrics(case, dtype=dt, residual_dtype=rdt)
                m.update(operands=op, residual=res)
                out["rows"].append(m)
                print(json.dumps(m), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
