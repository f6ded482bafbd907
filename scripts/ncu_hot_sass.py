"""Hot SASS instructions of one kernel from an ncu report captured with --import-source on.
    python scripts/ncu_hot_sass.py report.ncu-rep <kernel regex> [top N]     (no GPU needed)"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several launches may be concatenated: keep the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    if len(r) >= len(hdr) - 2:
        data.append(r)
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot_i = sum(int(r[ia]) for r in data); tot_s = sum(int(r[isamp]) for r in data)
print(f"{len(data)} SASS instructions, {tot_i} warp-instructions executed, {tot_s} stall samples")
idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]
print("line  samples%  exec%  sass")
for i in sorted(idx):
    r = data[i]
    print(f"{i:5d} {100*int(r[isamp])/max(tot_s,1):7.2f} {100*int(r[ia])/max(tot_i,1):6.2f}  {r[isrc].strip()[:100]}")
# coarse histogram over the program in 20 equal slices
n = len(data); sl = max(n // 20, 1)
print("slice: samples% exec%")
for k in range(0, n, sl):
    part = data[k:k + sl]
    print(f"  [{k:5d},{k+len(part):5d})  {100*sum(int(r[isamp]) for r in part)/max(tot_s,1):6.2f} {100*sum(int(r[ia]) for r in part)/max(tot_i,1):6.2f}")
