"""Executed instructions / stall samples per CUDA source line of one kernel: joins the SASS page of an ncu
report with `nvdisasm --print-line-info` of the library's cubin (same build), instruction by instruction.
usage: ncu_lines.py report.ncu-rep kernel_regex cubin mangled_name_substring units [top]"""
import csv, re, subprocess, sys, collections
rep, kern, cubin, mangled, units = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
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
iE, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
lines, cur = [], None
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print("sass in report", len(data), "sass in cubin", len(lines))
n = min(len(data), len(lines))
ex, st = collections.Counter(), collections.Counter()
for i in range(n):
    ex[lines[i]] += int(data[i][iE])
    st[lines[i]] += int(data[i][iS])
tot_e, tot_s = sum(ex.values()), sum(st.values())
print(f"thread-instr per unit {tot_e * 32 / units:.1f}")
src = {}
for (f, ln), _ in ex.most_common(top):
    try:
        if f not in src:
            import glob
            p = glob.glob(f"/root/repo/hash_join_codes_knl_b200/csrc/{f}")
            src[f] = open(p[0]).read().splitlines() if p else []
        text = src[f][ln - 1].strip()[:90] if src[f] else ""
    except Exception:
        text = ""
    print(f"{ex[(f, ln)] * 32 / units:6.2f} instr/unit  {st[(f, ln)] * 100 / max(1, tot_s):5.1f}% stalls  {f}:{ln}  {text}")
