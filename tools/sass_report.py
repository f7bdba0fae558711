"""Per-kernel counts of the Blackwell-native SASS mnemonics in libvpf.so -> profiles/sass_r02.txt
(the proof that the contraction kernels are tcgen05 / TMEM / TMA code, not recompiled mma.sync):

    UTCHMMA (tcgen05.mma, .2CTA = cta_group::2)   LDTM / STTM (tcgen05.ld / st, tensor memory)
    UTMALDG / UTMASTG (TMA tile load / store)     UTCBAR (tcgen05.commit -> mbarrier)
    HMMA (legacy warp-level mma.sync)             SYNCS (mbarrier ops)

Run by __graft_entry__.build() after the library is built (cuobjdump needs no GPU).
"""
import collections
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vitparticlefiltertracker_b200", "libvpf.so")
OUT = os.path.join(ROOT, "profiles", "sass_r02.txt")
MNEMONICS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF",
             "HMMA", "SYNCS", "FFMA2", "MUFU.EX2"]


def _tool(name):
    for cand in (shutil.which(name), f"/usr/local/cuda/bin/{name}"):
        if cand and os.path.exists(cand):
            return cand
    return None


def demangle(names):
    filt = _tool("cu++filt") or _tool("c++filt")
    if not filt or not names:
        return {n: n for n in names}
    out = subprocess.run([filt], input="\n".join(names), capture_output=True, text=True).stdout
    return dict(zip(names, out.splitlines()))


def short(name):
    depth = 0
    for i, ch in enumerate(name):                           # drop the parameter list: the first
        depth += ch == "<"                                  # "(" outside template brackets
        depth -= ch == ">"
        if ch == "(" and depth == 0:
            name = name[:i]
            break
    return name.replace("vpf::", "").replace("void ", "").replace("(int)", "")


def main(lib=LIB, out=OUT):
    cuobjdump = _tool("cuobjdump")
    if not cuobjdump or not os.path.exists(lib):
        print("sass_report: cuobjdump or libvpf.so missing, skipped", file=sys.stderr)
        return 1
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["instructions"] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                cur[mn] += 1
                if mn == "UTCHMMA.2CTA":
                    cur["UTCHMMA"] += 1
                break
    names = demangle(list(counts))
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA",
            "SYNCS", "FFMA2", "MUFU.EX2", "instructions"]
    rows = [(short(names[k]), [c[x] for x in cols]) for k, c in counts.items()]
    rows.sort(key=lambda r: (-r[1][0], -r[1][7], r[0]))
    width = min(86, max(len(r[0]) for r in rows))
    lines = ["# cuobjdump -sass vitparticlefiltertracker_b200/libvpf.so (sm_100a), per-kernel mnemonic counts",
             "# tools/sass_report.py; UTCHMMA = tcgen05.mma (.2CTA: cta_group::2), LDTM/STTM = tcgen05.ld/st,",
             "# UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync",
             f"{'kernel':<{width}} " + " ".join(f"{c:>12}" for c in cols)]
    for name, vals in rows:
        lines.append(f"{name[:width]:<{width}} " + " ".join(f"{v:>12}" for v in vals))
    tot = [sum(r[1][i] for r in rows) for i in range(len(cols))]
    lines.append(f"{'TOTAL (' + str(len(rows)) + ' kernels)':<{width}} " + " ".join(f"{v:>12}" for v in tot))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"sass_report: {len(rows)} kernels -> {out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
