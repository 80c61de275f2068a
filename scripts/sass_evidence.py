"""Per-kernel SASS evidence of the shipped library (no GPU needed): which kernels carry bulk-copy (TMA) instructions,
mbarriers, cluster barriers, FP64 tensor-core instructions and system-scope peer stores.

    python scripts/sass_evidence.py > profiles/r02b_sass_evidence.txt

Reads `cuobjdump -sass` of julia-phd-krylov-spdes_b200/libkrylov_b200.so (the sm_100a cubin the tests and bench load)."""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "julia-phd-krylov-spdes_b200", "libkrylov_b200.so")

# mnemonic class -> (regex on the instruction text, meaning)
CLASSES = OrderedDict([
    ("UBLKCP", (r"\bUBLKCP\b", "cp.async.bulk global->shared (TMA bulk copy)")),
    ("UBLKPF", (r"\bUBLKPF\b", "cp.async.bulk.prefetch.L2")),
    ("SYNCS", (r"\bSYNCS\b", "mbarrier arrive / try_wait / expect_tx")),
    ("UCGABAR", (r"\bUCGABAR", "cluster barrier (barrier.cluster)")),
    ("DMMA", (r"\bDMMA\b", "FP64 tensor-core mma.sync.m8n8k4.f64")),
    ("DFMA", (r"\bDFMA\b", "FP64 fused multiply-add")),
    ("ST.SYS", (r"\bST[G]?\.E\.64\.STRONG\.SYS\b", "relaxed system-scope 64-bit store (peer exchange push)")),
    ("LD.SYS", (r"\bLD[G]?\.E\.64\.STRONG\.SYS\b", "relaxed system-scope 64-bit load (peer exchange poll)")),
    ("SHFL", (r"\bSHFL\b", "warp shuffle (butterfly reductions)")),
    ("LDS", (r"\bLDS\b", "shared-memory load")),
])


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return out if len(out) == len(names) else names
    except Exception:
        return names


def short_name(p):
    p = p.replace("void ", "")
    if ">(" in p:
        return p[:p.index(">(") + 1].replace("(int)", "").replace("(bool)", "")
    return re.sub(r"\(.*", "", p)


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    funcs = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = {k: 0 for k in CLASSES}
            funcs[cur]["_n"] = 0
            funcs[cur]["_ex"] = {}
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        funcs[cur]["_n"] += 1
        for k, (rx, _) in CLASSES.items():
            if re.search(rx, ins):
                funcs[cur][k] += 1
                funcs[cur]["_ex"].setdefault(k, ins.strip())
    names = list(funcs)
    pretty = demangle(names)
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: architectures {arch}, {len(funcs)} kernels")
    print("# classes:")
    for k, (_, what) in CLASSES.items():
        print(f"#   {k:8s} {what}")
    print()
    hdr = "instr".rjust(7) + "".join(k.rjust(9) for k in CLASSES) + "  kernel"
    print(hdr)
    tot = {k: 0 for k in CLASSES}
    for n, p in zip(names, pretty):
        f = funcs[n]
        short = short_name(p)
        print(str(f["_n"]).rjust(7) + "".join(str(f[k]).rjust(9) for k in CLASSES) + "  " + short)
        for k in CLASSES:
            tot[k] += f[k]
    print("total".rjust(7) + "".join(str(tot[k]).rjust(9) for k in CLASSES))
    print()
    print("# one example instruction per class (first kernel that has it)")
    for k in CLASSES:
        for n, p in zip(names, pretty):
            if k in funcs[n]["_ex"]:
                short = short_name(p)
                print(f"{k:8s} {funcs[n]['_ex'][k]}    [{short}]")
                break
    return 0


if __name__ == "__main__":
    sys.exit(main())
