"""Aggregate an `ncu --page source --csv` SASS dump per CUDA source line using nvdisasm line info.

usage: ncu_source_agg.py <cubin> <kernel-mangled-substring> <ncu_sass.csv> [topN] [--outer]
Joins by instruction order (nvdisasm listing order == ncu SASS row order).  By default an instruction is
attributed to its innermost inlined source line; --outer attributes it to the line of b2pt_device.cuh's
`bounce` body (or the kernel) that the inline chain passes through, which groups whole stages.
"""
import collections
import os
import csv
import re
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
cubin, ksub, sass_csv = args[:3]
top = int(args[3]) if len(args) > 3 else 40
outer = "--outer" in sys.argv

txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
chains = []  # per instruction: list of (file, line) innermost first
infun = False
group, fresh = [], True
for ln in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        infun = ksub in m.group(1)
        group, fresh = [], True
        continue
    if not infun:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            group, fresh = [], False
        group.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        chains.append(list(group))
        fresh = True

rows = list(csv.reader(open(sass_csv)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
rows = rows[starts[-1]:]
print("kernel:", rows[0][1][:90], "(last of %d blocks)" % len(starts))
hdr = rows[1]
iE, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[2:] if len(r) > 10]
print("sass rows", len(data), "disasm instrs", len(chains))


def key_of(chain):
    if not chain:
        return None
    if not outer:
        return chain[0]
    # outermost location inside the device header's bounce()/trace() body, else the kernel line
    for loc in reversed(chain):
        if loc[0] == "b2pt_device.cuh":
            return loc
    return chain[-1]


agg = collections.defaultdict(lambda: [0, 0, 0])
totE = totT = totS = 0
for k, r in enumerate(data):
    key = key_of(chains[k]) if k < len(chains) else None
    e, t, s = int(r[iE]), int(r[iT]), int(r[iS])
    a = agg[key]
    a[0] += e
    a[1] += t
    a[2] += s
    totE += e
    totT += t
    totS += s
print("total warp-instr %d thread-instr %d avg active %.2f samples %d" % (totE, totT, totT / max(totE, 1), totS))
for key, (e, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s inst %6.2f%%  active %5.2f  stall-samples %5.2f%%" % (
        str(key), 100.0 * e / totE, t / max(e, 1), 100.0 * s / max(totS, 1)))

# ---- optional: attribute to the innermost NON-helper function (helpers = vector math lines < 100)
if "--func" in sys.argv:
    import bisect
    # function table of b2pt_device.cuh built from its own text: every line that starts a __device__ definition
    hdr_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "raytracingtherestofyourlife_b200", "csrc",
                            "b2pt_device.cuh")
    funcs = [(1, "header-top")]
    for n, line in enumerate(open(hdr_path).read().splitlines(), 1):
        m = re.match(r"^__device__[^(]*?(\w+)\(", line)
        if m:
            funcs.append((n, m.group(1)))
    helpers = {"dot3", "cross3", "mk3", "ld3", "operator", "rcp_fast", "pk2", "upk2", "fma2", "add2", "sub2", "rcp_safe",
               "normalize3", "length3", "mul3", "rmag3", "unit3", "denan3"}
    starts = [f[0] for f in funcs]
    agg2 = collections.defaultdict(lambda: [0, 0])
    for k, r in enumerate(data):
        chain = chains[k] if k < len(chains) else []
        name = "kernel/other"
        for loc in chain:
            if loc[0] == "b2pt_device.cuh":
                cand = funcs[bisect.bisect_right(starts, loc[1]) - 1][1]
                if cand in helpers or cand.startswith("operator"):
                    continue
                name = cand
                break
        agg2[name][0] += int(r[iE])
        agg2[name][1] += int(r[iT])
    print("---- by function (helpers attributed to their caller)")
    for name, (e, t) in sorted(agg2.items(), key=lambda kv: -kv[1][0]):
        print("%-26s inst %6.2f%%  active %5.2f" % (name, 100.0 * e / totE, t / max(e, 1)))
