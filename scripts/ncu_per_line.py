"""Attribute a kernel's executed warp-instructions to source lines: ncu's per-SASS-instruction counts (a `--set full
--import-source on` report) joined with the `nvdisasm -gi` line table of the same binary (built with -lineinfo).

    python scripts/ncu_per_line.py gpurun_out/prof_r1p.ncu-rep k_knn_scan superodom_b200/libsuperodom_b200.so [n_warps]

Prints warp-instructions per warp, share of instructions and of stall samples per innermost source line.  The report and the
library must come from the same build (the script checks that the opcodes line up).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def ncu_counts(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"] + os.environ.get("NCU_EXTRA", "").split(), capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    H = rows[hdr]
    ie, src, ss, ad = H.index("Instructions Executed"), H.index("Source"), H.index("# Samples"), H.index("Address")
    seen, U = set(), []
    for r in rows[hdr + 1:]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        if r[ad] in seen:                      # second launch of the same kernel: keep the first
            break
        seen.add(r[ad])
        U.append(r)
    toint = lambda x: int(x, 16) if x.startswith("0x") else int(x)
    base = toint(U[0][ad])
    return {toint(r[ad]) - base: (int(r[ie]), int(r[ss]) if r[ss].isdigit() else 0, r[src].strip()) for r in U}


def line_table(lib, kernel):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
        for f in sorted(os.listdir(d)):
            if not f.endswith(".cubin"):
                continue
            sass = subprocess.run(["nvdisasm", "-gi", os.path.join(d, f)], capture_output=True, text=True).stdout
            m, on, pending, fresh = {}, False, None, True
            for l in sass.splitlines():
                if l.startswith(".text."):
                    on = bool(re.search(kernel, l)) and not m
                    continue
                if on and l.startswith("\t.section"):
                    on = False
                if not on:
                    continue
                mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
                if mm:
                    if fresh:
                        pending, fresh = (mm.group(1), int(mm.group(2))), False
                    continue
                mm = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", l)
                if mm:
                    m[int(mm.group(1), 16)] = (pending, mm.group(2).strip())
                    fresh = True
            if m:
                return m
    return {}


def main():
    rep, kernel, lib = sys.argv[1:4]
    # SO_LIB_KERNEL: regex for the (mangled) section name in the library when the plain kernel name also matches another kernel
    cnt, m = ncu_counts(rep, kernel), line_table(lib, os.environ.get("SO_LIB_KERNEL", kernel))
    op = lambda t: (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
    bad = sum(1 for off, (_, _, t) in cnt.items() if off not in m or op(t) != op(m[off][1]))
    if bad:
        print(f"warning: {bad} of {len(cnt)} instructions do not line up -- report and library are not the same build", file=sys.stderr)
    per, smp = collections.Counter(), collections.Counter()
    for off, (c, s, _) in cnt.items():
        k = m[off][0] if off in m else None
        per[k] += c
        smp[k] += s
    tot, ts = sum(per.values()), max(sum(smp.values()), 1)
    n_warps = float(sys.argv[4]) if len(sys.argv) > 4 else max(c for c, _, _ in cnt.values())     # an instruction every warp runs once
    print(f"{kernel}: {tot / n_warps:.0f} warp-instructions per warp ({len(cnt)} SASS instructions)")
    cache = {}
    for k, c in per.most_common(60):
        if k is None:
            continue
        f, ln = k
        if f not in cache:
            try:
                cache[f] = open(f).read().split("\n")
            except OSError:
                cache[f] = []
        text = cache[f][ln - 1].strip()[:100] if ln <= len(cache[f]) else ""
        print(f"{c / n_warps:7.1f} {c / tot * 100:5.1f}% {smp[k] / ts * 100:5.1f}%  {os.path.basename(f)}:{ln}  {text}")


if __name__ == "__main__":
    main()
