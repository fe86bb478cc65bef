"""Summarise the SASS page of an ncu report: stall samples at synchronisation points and per code region.
    python tools/ncu_stalls.py report.ncu-rep [first_index last_index]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
acc = 0
keys = ("SYNCS", "LDTM", "STTM", "BAR.", "FENCE", "MEMBAR", "EXIT", "UTCBAR", "UTMA", "setmaxnreg", "USETMAXREG")
for i in range(lo, hi):
    r = data[i]
    n = int(r[isamp] or 0)
    acc += n
    if any(k in r[isrc] for k in keys) or n > tot / 150:
        st = {hdr[j]: int(r[j]) for j in range(hdr.index("stall_barrier"), hdr.index("stall_wait") + 1) if r[j] and int(r[j]) > 0}
        st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(i, "cum", acc, "n", n, "exec", r[iex], r[isrc][:70], st)
