"""Scan the PTX of kmm_tc.cu for reads of a tcgen05.ld destination register before the next tcgen05.wait::ld (the loads
are asynchronous; inline asm gives the compiler no dependency between the wait and the destinations).

    nvcc -gencode arch=compute_100a,code=compute_100a -O3 -std=c++17 -ptx -o /tmp/kmm_tc.ptx rlaopt_b200/csrc/kmm_tc.cu
    python scripts/ptx_tmem_ld_scan.py /tmp/kmm_tc.ptx
"""
import re
import sys

lines = open(sys.argv[1]).read().split("\n")
entry, pending, viol, nld = None, {}, {}, 0
for i, l in enumerate(lines):
    m = re.match(r"\s*(?:\.visible )?\.entry (\S+)\(", l)
    if m:
        entry, pending = m.group(1), {}
        continue
    t = l.strip()
    if not t or t.startswith("//") or t.startswith("."):
        continue
    if "tcgen05.ld" in t:
        mm = re.search(r"\{([^}]*)\}", t)
        for r in (mm.group(1).split(",") if mm else []):
            pending[r.strip()] = i
        nld += 1
        continue
    if "tcgen05.wait::ld" in t:
        pending = {}
        continue
    if pending:
        toks = re.findall(r"%[a-z]+\d+", t)
        if not toks:
            continue
        op = t.split()[0]
        reads = toks if op.startswith(("st.", "tcgen05.st", "mbarrier", "bar.", "@")) else toks[1:]
        for r in reads:
            if r in pending:
                viol.setdefault(entry, []).append((i + 1, t[:110], pending[r] + 1))
print(f"{nld} tcgen05.ld; kernels that read a destination before the next tcgen05.wait::ld: {len(viol)}")
for k, v in viol.items():
    print(k[:100], len(v), v[:3])
