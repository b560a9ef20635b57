"""Per-region instruction / stall summary of one kernel from an ncu report's source page.
usage: ncu_regions.py report.ncu-rep kernel_regex units_processed [block]"""
import csv, subprocess, sys, collections
rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
blk = int(sys.argv[4]) if len(sys.argv) > 4 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr, data = None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        if hdr is not None:
            break
        continue
    if hdr is None:
        hdr = r
        continue
    data.append(r)
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tots = sum(int(r[iSamp]) for r in data)
tot = sum(int(r[iE]) for r in data)
print("warp instructions", tot, "thread-instr per unit", tot * 32 / units, "sass lines", len(data))
for b in range(0, len(data), blk):
    e = sum(int(r[iE]) for r in data[b:b + blk]) * 32 / units
    s = sum(int(r[iSamp]) for r in data[b:b + blk]) * 100 / max(1, tots)
    ops = collections.Counter()
    for r in data[b:b + blk]:
        t = r[iS].strip().split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        ops[op.split(".")[0]] += int(r[iE])
    top = ", ".join(f"{k}:{v * 32 / units:.1f}" for k, v in ops.most_common(6))
    print(f"{b:5d} instr/unit {e:6.2f} stall% {s:5.1f}  {top}")
