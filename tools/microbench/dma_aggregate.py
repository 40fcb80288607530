"""Bare aggregate host<->device DMA bandwidth: N processes (one per GPU), pinned H2D and D2H running CONCURRENTLY on two streams,
no kernels.  This is the ceiling of any end-to-end pipeline on this box (bench.py's e2e moves 8 B per audio sample).

    python tools/microbench/dma_aggregate.py N [MB per direction per copy] [reps]

Prints one line per process and, from process 0, the aggregate (sum over processes of bytes / max over processes of time is
approximated by aligning the start on the wall clock and timing each process over the same number of copies)."""
import os
import subprocess
import sys
import time


def worker(gpu: int, mb: int, reps: int, both: int):
    os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu)
    import torch
    n = mb * (1 << 20) // 4
    hin = torch.empty(n, dtype=torch.float32, pin_memory=True).fill_(1.0)
    hout = torch.empty(n, dtype=torch.float32, pin_memory=True)
    din = torch.empty(n, dtype=torch.float32, device="cuda")
    dout = torch.ones(n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def go(k):
        for _ in range(k):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
    go(2)
    t = time.time(); time.sleep(3 - (t % 3))              # crude cross-process alignment
    t0 = time.perf_counter(); go(reps); dt = time.perf_counter() - t0
    gb = reps * n * 4 / 1e9
    print(f"gpu {gpu}: H2D {gb / dt:.1f} GB/s" + (f" + D2H {gb / dt:.1f} GB/s concurrently" if both else " (H2D only)"), flush=True)
    return gb * (2 if both else 1) / dt


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
        sys.exit(0)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    mb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    for both in (0, 1):
        for n in sorted({1, N} | {k for k in (2, 4) if k < N}):
            ps = [subprocess.Popen([sys.executable, __file__, "--worker", str(g), str(mb), str(reps), str(both)], stdout=subprocess.PIPE, text=True) for g in range(n)]
            outs = [p.communicate()[0] for p in ps]
            tot = 0.0
            for o in outs:
                for tok in o.split("\n"):
                    if "H2D" in tok:
                        tot += float(tok.split("H2D")[1].split()[0]) * (2 if both else 1)
            print(f"n_gpus={n} mode={'H2D+D2H' if both else 'H2D'}: aggregate {tot:.1f} GB/s ({tot / n:.1f} per GPU)", flush=True)
