"""Per-kernel SASS instruction evidence of the in-tree library: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "photo-slam_b200/lib/libpsb200.so"
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
pats = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("REDG", r"\bRED(G)?\b|\bRED\."), ("ATOM", r"\bATOM[SG]?\b|\bATOM\."), ("SHFL", r"\bSHFL"), ("MUFU", r"\bMUFU"),
        ("LDG", r"\bLDG"), ("STG", r"\bSTG|\bST\.E"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("MEMBAR", r"\bMEMBAR"), ("FFMA", r"\bFFMA")]
rows, cur = [], None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("psb::", "")
        name = re.sub(r"<\(int\)(\d+)>", r"<\1>", re.sub(r"<\(bool\)(\d+)>", r"<\1>", name)).split("(")[0]
        cur = {"name": name, "total": 0, **{k: 0 for k, _ in pats}}
        rows.append(cur)
    elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        cur["total"] += 1
        for k, p in pats:
            if re.search(p, line):
                cur[k] += 1
print(f"# cuobjdump -sass {LIB} (sm_100a): static SASS instruction counts per kernel (round 2 final kernels)")
print("# UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier arrive/try_wait, REDG = global reductions (RED.E.ADD...), MEMBAR = fences")
print("kernel | total | " + " | ".join(k for k, _ in pats))
for r in sorted(rows, key=lambda r: r["name"]):
    print(f"{r['name']} | {r['total']} | " + " | ".join(str(r[k]) for k, _ in pats))
