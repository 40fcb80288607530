"""Per-kernel stall-sample breakdown from the ncu source page.  Usage: python tools/ncu_stalls.py rep kernel_regex [launch_skip]"""
import csv, collections, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:90])
hdr = rows[1]; I = {n: i for i, n in enumerate(hdr)}
seen, data = set(), []
for r in rows[2:]:
    if len(r) >= len(hdr) - 2 and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); data.append(r)
tot = sum(int(r[I["# Samples"]]) for r in data)
names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n[6:]: sum(int(r[I[n]]) for r in data) for n in names}
print("samples", tot, {k: f"{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.01 * tot})
def opc(src):
    t = src.split(); return t[1] if t[0].startswith("@") else t[0]
byop, execd = collections.Counter(), collections.Counter()
for r in data:
    byop[opc(r[I["Source"]])] += int(r[I["# Samples"]]); execd[opc(r[I["Source"]])] += int(r[I["Instructions Executed"]])
te = sum(execd.values())
print("by opcode (samples%, executed%):", [(o, f"{100*s/tot:.1f}", f"{100*execd[o]/te:.1f}") for o, s in byop.most_common(9)])
wf = sum(int(r[I["L1 Wavefronts Shared"]]) for r in data); xwf = sum(int(r[I["L1 Wavefronts Shared Excessive"]]) for r in data)
print("smem wavefronts", wf, "excessive", xwf, f"({100*xwf/max(1,wf):.1f}%)", " warp-instr executed", te)
for r in sorted(data, key=lambda r: -int(r[I["# Samples"]]))[:10]:
    st = {n[6:]: int(r[I[n]]) for n in names if int(r[I[n]]) > 0}
    print(f"{100*int(r[I['# Samples']])/tot:5.1f}% ", r[I["Source"]].strip()[:56].ljust(56), dict(sorted(st.items(), key=lambda x: -x[1])[:2]))
