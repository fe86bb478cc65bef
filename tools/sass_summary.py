"""Per-kernel SASS evidence: counts of the tcgen05 / TMEM / TMA instructions in every kernel of libtransvae_sm100.so
(`cuobjdump -sass`; runs without a GPU).  UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG /
UTMAREDG = TMA load / store / reduce, UTCBAR = tcgen05.commit.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "deepl-project_b200", "libtransvae_sm100.so")
MNEM = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "MUFU.EX2", "MUFU.TANH", "FFMA2", "HFMA2",
        "ATOMS", "RED", "ATOMG"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts, total, name = collections.OrderedDict(), {}, None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(.*", "", name)
        counts[name] = collections.Counter()
        total[name] = 0
        continue
    if name is None or "/*" not in ln:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not m:
        continue
    op = m.group(1)
    total[name] += 1
    for k in MNEM:
        if op == k or op.startswith(k + ".") or (k.startswith("MUFU") and op.startswith(k)):
            counts[name][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: instruction counts per kernel (sm_100a)")
print(f"{'kernel':90s} {'instr':>7s} " + " ".join(f"{k:>9s}" for k in MNEM))
tot = collections.Counter()
for k, c in sorted(counts.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + total[kv[0]])):
    print(f"{k[:90]:90s} {total[k]:7d} " + " ".join(f"{c[m]:9d}" for m in MNEM))
    tot.update(c)
print(f"{'TOTAL':90s} {sum(total.values()):7d} " + " ".join(f"{tot[m]:9d}" for m in MNEM))
