"""Print the hot SASS lines (stall samples) of a kernel from `ncu --page source --csv` output."""
import csv, subprocess, sys
rep = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]; ia = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
data = [r for r in rows[hi + 1:end] if len(r) > isamp and r[isamp].strip().isdigit()]
tot = sum(int(r[isamp]) for r in data)
print("lines", len(data), "samples", tot, "warp-instr", sum(int(r[iex]) for r in data))
keys = ("SYNCS", "UTMALDG", "UTCHMMA", "BAR", "UTCBAR", "LDTM", "STTM", "MUFU", "EXIT", "UBLKCP")
acc = 0
for k, r in enumerate(data):
    s = int(r[isamp]); acc += s
    if s >= thr or any(x in r[ia] for x in keys):
        print(f"{k:5d} cum{100*acc/tot:5.1f}% {s:5d} {r[iex]:>9s}  {r[ia][:100]}")
