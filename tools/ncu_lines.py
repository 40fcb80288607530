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
si, wi, xi, ii = H.index("# Samples"), H.index("L1 Wavefronts Shared"), H.index("L1 Wavefronts Shared Excessive"), H.index("Instructions Executed")
stall_cols = [(n[6:], i) for i, n in enumerate(H) if n.startswith("stall_") and "Not Issued" not in n]
lines = [r for r in rows[hi + 1:] if r and r[0].strip().isdigit()]
tot = sum(int(r[si]) for r in lines); twf = sum(int(r[wi]) for r in lines)
print(rows[1][1][:100]); print("total samples", tot, "smem wavefronts", twf)
for r in sorted(lines, key=lambda r: -int(r[si]))[:top]:
    st = sorted(((n, int(r[i])) for n, i in stall_cols if int(r[i]) > 0), key=lambda x: -x[1])[:3]
    print(f"{100*int(r[si])/tot:5.1f}%  wf {100*int(r[wi])/max(1,twf):4.1f}% (x{int(r[xi])/max(1,int(r[wi])):.2f})  inst {int(r[ii]):>11d}  L{r[0]:>4s}  {r[1].strip()[:84]:84s} {st}")
