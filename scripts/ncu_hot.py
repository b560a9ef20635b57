"""Summarise an ncu report's source page: top SASS instructions by executed count / stall samples."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout.splitlines()
blocks, cur = [], None
for row in csv.reader(out):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
b = blocks[0]
h = b["hdr"]
iS, iE, iSamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
rows = b["rows"]
tot_e = sum(int(r[iE]) for r in rows); tot_s = sum(int(r[iSamp]) for r in rows)
print(b["name"][:80], "instructions", tot_e, "samples", tot_s, "sass lines", len(rows))
print("--- by stall samples")
for r in sorted(rows, key=lambda r: -int(r[iSamp]))[:top]:
    print(f"{int(r[iSamp])*100/tot_s:5.1f}%  exec {int(r[iE])*100/tot_e:5.1f}%  #{rows.index(r):4d} {r[iS].strip()[:90]}")
