"""Per-CUDA-source-line stall samples from an .ncu-rep.  Usage: python tools/ncu_lines.py rep kernel_regex [launch_skip] [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hi]
def col(name):
    return H.index(name) if name in H else None
si, wi, xi, ii = col("# Samples"), col("L1 Wavefronts Shared"), col("L1 Wavefronts Shared Excessive"), col("Instructions Executed")
def geti(r, i):
    return int(r[i]) if i is not None and r[i].strip().isdigit() else 0
stall_cols = [(n[6:], i) for i, n in enumerate(H) if n.startswith("stall_") and "Not Issued" not in n]
lines = [r for r in rows[hi + 1:] if r and r[0].strip().isdigit()]
tot = sum(geti(r, si) for r in lines); twf = sum(geti(r, wi) for r in lines)
print(rows[1][1][:100]); print("total samples", tot, "smem wavefronts", twf)
for r in sorted(lines, key=lambda r: -geti(r, si))[:top]:
    st = sorted(((n, geti(r, i)) for n, i in stall_cols if geti(r, i) > 0), key=lambda x: -x[1])[:3]
    print(f"{100*geti(r, si)/max(1,tot):5.1f}%  wf {100*geti(r, wi)/max(1,twf):4.1f}% (x{geti(r, xi)/max(1,geti(r, wi)):.2f})  inst {geti(r, ii):>11d}  L{r[0]:>4s}  {r[1].strip()[:84]:84s} {st}")
