"""List the instructions with the most warp-stall samples of one kernel in an .ncu-rep (source page, SASS view).
usage: python tools/ncu_source_hot.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[k]
ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for r in rows[k + 1:]:
    if len(r) <= iex or r[0] == "Address" or r[0] == "Kernel Name":
        break
    data.append((r[ia].strip(), int(r[isamp] or 0), int(r[iex] or 0)))
tot = sum(d[1] for d in data) or 1
print(f"{rows[0][1][:80]}: {len(data)} instructions, {tot} stall samples, {sum(d[2] for d in data)} warp instructions executed")
for idx, (s, n, ex) in sorted(enumerate(data), key=lambda kv: -kv[1][1])[:top]:
    print(f"{idx:5d} {n:7d} {100 * n / tot:5.1f}%  ex={ex:9d}  {s[:100]}")
