"""Stall samples per SASS instruction of one kernel, by reason: the instructions that hold most samples, with the
memory wavefront columns.  usage: ncu_stalls.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
by_reason = {h: sum(int(r[ix[h]]) for r in data) for h in reasons}
print("samples", tot, {k[6:]: round(100 * v / tot, 1) for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1]) if v * 50 > tot})
for n, r in sorted(enumerate(data), key=lambda nr: -int(nr[1][ix["# Samples"]]))[:top]:
    s = int(r[ix["# Samples"]])
    rs = sorted(((int(r[ix[h]]), h[6:]) for h in reasons), reverse=True)[:2]
    print(f"{n:5d} {100*s/tot:5.1f}%  {r[ix['Source']].strip()[:58]:58s} {rs[0][1]}={rs[0][0]} {rs[1][1]}={rs[1][0]}  wf={r[ix['L1 Wavefronts Shared']]}")
