"""Per-kernel counts of the SASS instructions that prove a Blackwell-native kernel (tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM / STTM,
TMA bulk copies -> UBLKCP / UTMA*, legacy tensor path -> HMMA) in the in-tree libaqgnn.so:
    python scripts/sass_summary.py > profiles/sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "alphaquoridorgnn_b200", "libaqgnn.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "HMMA", "ACQBULK", "VHMNMX"]
name, rows = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        rows[name] = dict.fromkeys(pats, 0)
        rows[name]["instructions"] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(\S.*?);", line)
    if m and name:
        rows[name]["instructions"] += 1
        for p in pats:
            if re.search(r"\b" + p, m.group(1)):
                rows[name][p] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS instruction counts per kernel (cuobjdump -sass; sm_100a)")
print("kernel".ljust(64) + "instr".rjust(7) + "".join(p.rjust(9) for p in pats))
for k, v in sorted(rows.items(), key=lambda kv: -kv[1]["UTCHMMA"]):
    print(k[:63].ljust(64) + str(v["instructions"]).rjust(7) + "".join(str(v[p]).rjust(9) for p in pats))
