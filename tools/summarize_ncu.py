"""Turns the raw ncu artefacts brought back in gpurun_out/ into the committed summaries under
profiles/ (run here, no GPU needed):

    python tools/summarize_ncu.py gpurun_out/launches.csv gpurun_out/prof_gemm.ncu-rep
"""
import collections
import csv
import json
import subprocess
import sys

M, D = 4096 * 197, 768
ALG = {"qkv": (M * D * 2 + 3 * D * D * 2 + M * 3 * D * 2, 2.0 * M * 3 * D * D),
       "proj": (M * D * 2 + D * D * 2 + 2 * M * D * 4, 2.0 * M * D * D),
       "fc1": (M * D * 2 + 4 * D * D * 2 + M * 4 * D * 2, 2.0 * M * 4 * D * D),
       "fc2": (M * 4 * D * 2 + 4 * D * D * 2 + 2 * M * D * 4, 2.0 * M * 4 * D * D)}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def launches(path, tag="r01", head=None):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        agg[r[ki][:64]][0] += 1
        agg[r[ki][:64]][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    lines = [f"{k:64s} n={v[0]:4d} total={v[1] / 1e6:9.3f} ms share={v[1] / tot * 100:5.1f}% "
             f"avg={v[1] / v[0] / 1e3:9.1f} us" for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    head = head or ("ncu --metrics gpu__time_duration.sum --clock-control none -s 385 -c 300, bench.py --steps 2 "
                    "--warmup 3 (1xB200, ViT-B/16, N=4096)\nper-launch times are cold-cache and serialised at "
                    "full clocks: compare SHARES with bench.py's share_of_step\n\n")
    open(f"profiles/launches_{tag}_summary.txt", "w").write(head + "\n".join(lines) + "\n")
    open(f"profiles/launches_{tag}.csv", "w").write(open(path).read())
    print("\n".join(lines[:8]))


def gemm(path):
    out = subprocess.run(f"ncu -i {path} --page raw --csv", shell=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    g = lambda r, n: float(r[col[n]].replace(",", ""))
    res = []
    for r, shape in zip(rows[2:], ["qkv", "proj", "fc1", "fc2", "qkv"]):
        rd = g(r, "dram__bytes_read.sum") * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = g(r, "dram__bytes_write.sum") * UNIT[units[col["dram__bytes_write.sum"]]]
        res.append({"kernel": r[col["Kernel Name"]][:64], "shape": shape,
                    "duration_ms_under_ncu": g(r, "gpu__time_duration.sum") * TIME[units[col["gpu__time_duration.sum"]]],
                    "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
                    "algorithmic_bytes": ALG[shape][0], "flops": ALG[shape][1],
                    "tensor_pipe_active_pct": g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "sm_clock_ghz": g(r, "sm__cycles_elapsed.avg.per_second"),
                    "dram_throughput_pct": g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    "l2_hit_pct": g(r, "lts__t_sector_hit_rate.pct"),
                    "issue_active_pct": g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "registers": g(r, "launch__registers_per_thread")})
    fam = {x["shape"]: x for x in res[:4]}
    summary = {"source": f"ncu --set full --clock-control none, bench.py --steps 2 --warmup 3 (1xB200, ViT-B/16, "
                         f"N=4096); {path}",
               "launches": res,
               "gemm_family_traffic_bytes_per_block": sum(v["traffic_bytes"] for v in fam.values()),
               "gemm_family_algorithmic_bytes_per_block": sum(v["algorithmic_bytes"] for v in fam.values())}
    json.dump(summary, open("profiles/ncu_gemm_r01.json", "w"), indent=1)
    for x in res:
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in x.items()
               if k in ("shape", "duration_ms_under_ncu", "traffic_bytes", "tensor_pipe_active_pct", "sm_clock_ghz",
                        "dram_throughput_pct", "registers")})


KEYS = {
    "duration_ms": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_clock_ghz": "sm__cycles_elapsed.avg.per_second",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "warp_instructions": "smsp__inst_executed.sum",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size", "block": "launch__block_size",
    "ctas_per_sm_limit_smem": "launch__occupancy_limit_shared_mem",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
}


def kernels(path, out, note=""):
    """Generic per-launch summary of an `ncu --set full` report -> profiles/<out>.json, with the
    top stall reasons, stamped with the git commit it was summarised at."""
    raw = subprocess.run(f"ncu -i {path} --page raw --csv", shell=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        if name not in col or r[col[name]] in ("", "n/a"):
            return None
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]]
        if u in UNIT:
            v *= UNIT[u]
        if name == "gpu__time_duration.sum":
            v *= TIME.get(u, 1.0)
        if u.lower() == "ghz" or name.endswith("per_second") and u == "Ghz":
            pass
        return v
    stall_cols = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[col["Kernel Name"]][:90]}
        d.update({k: val(r, n) for k, n in KEYS.items()})
        st = sorted(((float(r[col[h]].replace(",", "")), h) for h in stall_cols
                     if r[col[h]] not in ("", "n/a")), reverse=True)[:5]
        d["top_stalls_per_issue"] = {h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): round(v, 2)
                                     for v, h in st}
        res.append(d)
    sha = subprocess.run("git rev-parse --short HEAD", shell=True, capture_output=True, text=True).stdout.strip()
    json.dump({"source": f"ncu --set full --clock-control none --import-source on; {path}", "note": note,
               "summarised_at_git": sha, "launches": res}, open(f"profiles/{out}.json", "w"), indent=1)
    for d in res:
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()})


def block(path, tag, note=""):
    """`ncu --set full` of tools/prof_kernels.py from the crop kernel of step 1 on (crop, patch embed,
    rowstats, qkv, attention, proj, fc1, fc2, qkv ...) -> profiles/ncu_block_<tag>.json (every launch)
    and profiles/ncu_gemm_<tag>.json (the GEMM family's DRAM traffic in bench.py's `roofline.traffic`
    format, stamped with the hash of the GEMM sources and the commit)."""
    import hashlib
    kernels(path, f"ncu_block_{tag}", note)
    launches_ = json.load(open(f"profiles/ncu_block_{tag}.json"))["launches"]
    esz = 2                                               # fp16 residual stream
    alg = {"qkv": M * D * 2 + 3 * D * D * 2 + M * 3 * D * 2, "proj": M * D * 2 + D * D * 2 + 2 * M * D * esz,
           "fc1": M * D * 2 + 4 * D * D * 2 + M * 4 * D * 2, "fc2": M * 4 * D * 2 + 4 * D * D * 2 + 2 * M * D * esz}
    mode_of = {"qkv": ("<256, 5, 2>", "<256, 0, 2>"), "proj": ("<256, 4, 2>",), "fc1": ("<256, 6, 2>", "<256, 1, 2>"),
               "fc2": ("<256, 2, 2>",)}
    res = []
    for shape, modes in mode_of.items():
        x = next(l for l in launches_ if any(m in l["kernel"] for m in modes))
        res.append({"shape": shape, "kernel": x["kernel"][:64], "duration_ms_under_ncu": x["duration_ms"],
                    "dram_read_bytes": x["dram_read_bytes"], "dram_write_bytes": x["dram_write_bytes"],
                    "traffic_bytes": x["dram_read_bytes"] + x["dram_write_bytes"], "algorithmic_bytes": alg[shape],
                    "flops": ALG[shape][1], "tensor_pipe_active_pct": x["tensor_pipe_active_pct"],
                    "sm_clock_ghz": x["sm_clock_ghz"], "dram_throughput_pct": x["dram_throughput_pct"],
                    "l2_hit_pct": x["l2_hit_pct"], "issue_active_pct": x["issue_active_pct"],
                    "registers": x["registers"]})
    h = hashlib.sha256()
    for n in ("gemm.cuh", "sm100.cuh"):
        h.update(open("vitparticlefiltertracker_b200/csrc/" + n, "rb").read())
    sha = subprocess.run("git rev-parse --short HEAD", shell=True, capture_output=True, text=True).stdout.strip()
    json.dump({"source": f"ncu --set full --clock-control none --import-source on, tools/prof_kernels.py (1xB200, "
                         f"ViT-B/16, N=4096, fp16 residual stream, LayerNorm folded into qkv / fc1); {path}",
               "residual": "fp16", "launches": res,
               "gemm_family_traffic_bytes_per_block": sum(r["traffic_bytes"] for r in res),
               "gemm_family_algorithmic_bytes_per_block": sum(alg.values()),
               "source_sha16": h.hexdigest()[:16], "git_sha": sha},
              open(f"profiles/ncu_gemm_{tag}.json", "w"), indent=1)
    for r in res:
        print(r["shape"], round(r["traffic_bytes"] / 1e9, 2), "GB vs", round(r["algorithmic_bytes"] / 1e9, 2),
              "tensor %", round(r["tensor_pipe_active_pct"], 1), "GHz", round(r["sm_clock_ghz"], 3))


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[1] == "--block":    # --block report.ncu-rep tag [note]
        block(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
        sys.exit(0)
    if len(sys.argv) > 3 and sys.argv[1] == "--kernels":  # --kernels report.ncu-rep out_name [note]
        kernels(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
        sys.exit(0)
    if len(sys.argv) > 3 and sys.argv[2] == "--tag":     # launch list of another workload
        launches(sys.argv[1], sys.argv[3], sys.argv[4] + "\n\n" if len(sys.argv) > 4 else None)
        sys.exit(0)
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        gemm(sys.argv[2])
