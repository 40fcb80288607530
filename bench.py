#!/usr/bin/env python
"""Benchmark of the GTCRN-Micro inference hot path (BASELINE.json: audio-seconds per second per B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the whole hot path (STFT -> model -> iSTFT) over one batch of 1024 synthetic
10 s / 16 kHz noisy utterances per GPU (BASELINE.json configs[1]); utterances are independent, so N GPUs
process N such batches with no collective (weak scaling).  One JSON line is printed by rank 0.

  value         device-timed (CUDA events on the launching stream), inputs resident in HBM
  e2e           same metric through `GTCRNMicro.enhance_host` (C ABI gtcrnb_enhance_host): pinned HOST
                buffers in, pinned HOST buffers out, H2D / D2H inside the timed region
  roofline      dominant kernel class (by time), timed live with CUDA events inside the library; HBM roof for the tensor-core kernels
  cpu_baseline  the reference's own CPU implementation (baseline/_ref; oracle port if absent) on this box's host
                cores: offline audio-s/s and, per BASELINE.md section 3, max real-time streams of its frame-by-frame model
  streaming     frame-by-frame path (BASELINE configs[2]): p50/p99/max hop latency over 625 hops for 4,096 / 16,384
                streams, wav in / wav out through pinned host buffers (copies inside the hop), HBM roofline of the hop,
                and the largest stream count whose worst hop meets the 16 ms deadline (searched, then confirmed)

  parity        a fixed 16-utterance subset of the TIMED batch's output compared with the oracle (checker only,
                outside the timed region): waveform max-abs <= 1e-4 and SI-SNR delta <= 0.01 dB or the run fails

`--impl reference` times the reference's own CPU implementation of the same path: the unmodified bglid/GTCRN-Micro
package installed into the git-ignored baseline/_ref by oracle/install_reference.sh (kind "reference": its
GTCRNMicro between torch.stft / torch.istft exactly as infer.py:60-76, its StreamGTCRNMicro + convert_to_stream for
the streaming figure).  When baseline/_ref is absent the arm falls back to the oracle port (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"
SR = 16000
SECONDS = 10
L_UTT = SR * SECONDS                      # 160,000 samples -> T = 626 frames
FRAMES_PER_UTT = 1 + L_UTT // 256
MAC_PER_FRAME = 614_530                   # conv/linear MACs of the reference's own ops
DEC_GT_MAC_PER_FRAME = 84_582             # one decoder GTConvBlock
FP32_PEAK_TFLOPS = 73.4                   # measured on this pool's B200: tools/microbench/ffma_feed.cu -> profiles/ffma_feed_r1.txt
# Per kernel class of the GTCRN-Micro path: algorithmic MACs and HBM bytes per 16 ms frame,
# the roof that bounds the FFMA build of the class, and ncu DRAM bytes per frame where a capture exists (profiles/).
A33 = 33 * 16 * 4                          # one [33][16] fp32 activation frame: 2,112 B
MICRO_CLASS_INFO = {
    "front": {"kernel": "k_front<FROM_WAV> (STFT, ERB, SFE-Lite, en0, en1)", "mac": 36_864 + 1_161 + 15_600 + 42_240, "alg_bytes": 1024 + 2056 + 65 * 64 + A33, "bound": "fp32_fma",
              "why": "FP32 FMA / shared-memory bound (radix-4 FFTs and two strided convs on the CUDA cores)"},
    "enc_gtconv": {"kernel": "k_gtconv<ENC> (encoder GTConvBlock)", "mac": 13_288, "alg_bytes": 2 * A33, "bound": "hbm",
                   "why": "13.3 k MAC per 4,224 B: HBM / latency bound", "dram_bytes": 4212, "dram_note": "ncu r1_v13: 2.70 GB per launch of 641,024 frames"},
    "gtcn": {"kernel": "k_gtcn (4 dilated TCNs, FFMA)", "mac": 73_920, "alg_bytes": 2 * A33, "bound": "fp32_fma",
             "why": "FFMA build: FMA pipe 66 % busy (ncu r1_v13)", "dram_bytes": 4150, "dram_note": "ncu r1_v13: 2.66 GB per launch"},
    "dec_gtconv": {"kernel": "k_gtconv<DEC> (decoder GTConvBlock, FFMA)", "mac": 84_582, "alg_bytes": 3 * A33, "bound": "fp32_fma",
                   "why": "FFMA build: FMA pipe 70 % busy (ncu r1_v13)", "dram_bytes": 8555,
                   "dram_note": "ncu r1_v13: the untouched channel half is re-read in the store phase"},
    "back": {"kernel": "k_back<TO_WAV> (de3, de4, ERB expand, mask, iSTFT, overlap-add)", "mac": 42_240 + 10_400 + 24_576, "alg_bytes": A33 + A33 + 65 * 64 + 2056 + 1024, "bound": "fp32_fma",
             "why": "FP32 FMA / shared-memory bound"},
}
# the same classes when they run on the tensor cores (GTCRNB_TC bits, default on): the contraction leaves the FMA pipe, what remains is
# bounded by HBM (and, inside the SM, by the L1TEX data pipe that carries shared-memory, global and tensor-memory traffic: ncu r2)
MICRO_TC_INFO = {
    "dec_gtconv": {"kernel": "k_gtconv_dec_tc (decoder GTConvBlock, 3x3 conv as tcgen05 kind::f16 MMAs on fp16 hi/lo operands in tensor memory)", "bound": "hbm",
                   "why": "the 2,304-MAC 3x3 contraction runs on the tensor cores (9 MMAs per 99-position tile); the rest is bounded by HBM traffic "
                          "(6,336 B per frame) and the dependent chain of a step: ncu r2 v27 DRAM 40 %, issue 41 %, tensor pipe < 10 %; clock64 trace of a step in DESIGN 4.1a",
                   "dram_bytes": 6396, "dram_note": "ncu r2 (profiles/ncu_r2_v30_summary.txt): 2.78 GB read + 1.32 GB written per launch of 641,024 frames = 1.01 x algorithmic"},
    "gtcn": {"kernel": "k_gtcn_tc (4 dilated TCNs, pointwise convs as tcgen05 kind::tf32 3xTF32 MMAs, A operand in tensor memory)", "bound": "hbm",
             "why": "the 512 of 560 MACs per position-TCN that are pointwise contractions run on the tensor cores; the kernel is bound by "
                    "the dependent chain of a TCN (clock64 trace: 4,450 cycles per 32-frame tile and TCN with 6 CTAs per SM; ncu r2 v27 issue 55 %; "
                    "cutting 7 % of its instructions changed nothing), nearest hardware roof HBM", "dram_bytes": 4155,
             "dram_note": "ncu r2 (profiles/ncu_r2_v30_summary.txt): 1.35 GB read + 1.31 GB written per launch"},
}
WEIGHTS = os.path.join(ROOT, "tests", "golden", "weights_dns3.npz")
# BASELINE configs[3], `--variant full`: the original-GTCRN-style architecture (parity unpinned, oracle/gtcrn_full_oracle.py).
# conv / linear / GRU MACs per frame of its own ops: ERB 61,440 + en0 46,800 + en1 21,120 + 6 GTConv x 23,456 (1x1 24->16, dw 3x3,
# 1x1 16->8, TRA GRU 8->16 + fc) + 2 DPGRNN x 61,248 (intra 4 chains x 33 x 144 + fc, inter 66 chains x 384 + fc) + de3 21,120 +
# de4 10,400 + ERB^-1 24,576 + mask 1,028
FULL_MAC_PER_FRAME = 449_716
FULL_CLASS_MAC = {"front": 61_440 + 46_800 + 21_120, "full_enc_gtconv": 23_456, "full_dec_gtconv": 23_456, "dp_intra": 19_008 + 8_448,
                  "dp_inter": 25_344 + 8_448, "back": 21_120 + 10_400 + 24_576 + 1_028}
FULL_SEED = 20261017


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def synth_on_device(B: int, L: int, device, seed: int, keep_clean=None):
    """Same recipe as oracle.synth_batch generated on the device in fp32: 5 harmonics of
    f0~U(90,250) Hz, 4 Hz syllable envelope, RMS 0.05, white noise at SNR~U(0,15) dB, clipped."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty(B, L, device=device, dtype=torch.float32)
    t = torch.arange(L, device=device, dtype=torch.float32) / SR
    for b0 in range(0, B, 128):
        nb = min(128, B - b0)
        f0 = 90 + 160 * torch.rand(nb, 1, generator=g, device=device)
        clean = torch.zeros(nb, L, device=device)
        for h in range(5):
            ph = 6.2831853 * torch.rand(nb, 1, generator=g, device=device)
            clean += torch.sin(6.2831853 * f0 * (h + 1) * t + ph) / (h + 1)
        env = 0.5 * (1 + torch.sin(6.2831853 * 4.0 * t + 6.2831853 * torch.rand(nb, 1, generator=g, device=device)))
        clean *= env
        clean *= 0.05 / clean.pow(2).mean(dim=1, keepdim=True).sqrt()
        snr = 15 * torch.rand(nb, 1, generator=g, device=device)
        noise = torch.randn(nb, L, generator=g, device=device) * (0.05 / 10 ** (snr / 20))
        out[b0:b0 + nb] = (clean + noise).clamp_(-1, 1)
        if keep_clean is not None:
            for j, i in enumerate(keep_clean[0]):
                if b0 <= i < b0 + nb:
                    keep_clean[1][j] = clean[i - b0].cpu()
    return out


class ClockSampler:
    """SM clock / throttle reasons sampled every ~5 ms DURING the timed region (NVML in a thread; nvidia-smi -lms as fallback)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
               ("hw_power_brake", 0x80))

    def __init__(self, index: int):
        import threading
        self.index, self.samples, self.reasons, self.power = index, [], set(), []
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # honour CUDA_VISIBLE_DEVICES remapping: match by PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hh).bus == bus:
                        self.h = hh
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is None:
            return
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz, "power_w_max": max(self.power) if self.power else None,
                "samples": len(sm), "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference on the host cores
# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


class CpuReference:
    """The reference's CPU implementation of the path.  kind "reference": the unmodified package from baseline/_ref
    (oracle/install_reference.sh) driven exactly like infer.py:60-76 and gtcrn_micro_stream.py:618-636; kind "port": the
    oracle restatement (same ATen ops), used only when baseline/_ref is missing."""

    def __init__(self, variant: str = "micro", weights=None):
        import gtcrn_oracle as O
        self.O = O
        self.variant = variant
        self.kind = "port"
        self.model = self.smodel = None
        if variant == "full":
            # no reference implementation of this variant exists (only commented-out lines in gtcrn_micro.py): the CPU arm is the
            # builder-written oracle (parity unpinned)
            import gtcrn_full_oracle as FO
            self.FO = FO
            self.W = weights if weights is not None else full_variant_weights()
            return
        self.W = O.load_weights_npz(WEIGHTS)
        if os.path.isdir(os.path.join(REF_DIR, "gtcrn_micro")):
            try:
                import types
                if REF_DIR not in sys.path:
                    sys.path.insert(0, REF_DIR)
                sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))   # gtcrn_micro_stream.py:8 imports it; only __main__ uses it
                from gtcrn_micro.models.gtcrn_micro import GTCRNMicro
                from gtcrn_micro.streaming.conversion.convert import convert_to_stream
                from gtcrn_micro.streaming.gtcrn_micro_stream import StreamGTCRNMicro
                ck = os.path.join(REF_DIR, "gtcrn_micro", "ckpts", "best_model_dns3.tar")
                sd = torch.load(ck, map_location="cpu", weights_only=True)["model"] if os.path.exists(ck) else self.W
                self.model = GTCRNMicro().eval()
                self.model.load_state_dict(sd)
                self.smodel = StreamGTCRNMicro().eval()
                convert_to_stream(self.smodel, self.model)
                self.kind = "reference"
            except Exception as e:      # a broken install must not sink the bench: fall back to the port and say so
                sys.stderr.write(f"bench: reference package in {REF_DIR} unusable ({type(e).__name__}: {e}); using the oracle port\n")
                self.model = self.smodel = None

    def enhance(self, x: torch.Tensor) -> torch.Tensor:
        if self.variant == "full":
            return self.FO.enhance(self.W, x)
        if self.model is None:
            return self.O.enhance(self.W, x)
        with torch.inference_mode():                                               # infer.py:60-76, batched
            win = torch.hann_window(512).pow(0.5)
            spec = torch.stft(x, 512, 256, 512, win, return_complex=False)
            out = self.model(spec)
            return torch.istft(torch.view_as_complex(out.contiguous()), 512, 256, 512, win)

    def stream_hop_seconds(self, B: int, warm: int = 3, hops: int = 10) -> float:
        """Mean seconds per hop of the frame-by-frame model batched over B streams."""
        g = torch.Generator().manual_seed(B)
        frames = [torch.randn(B, 257, 1, 2, generator=g) for _ in range(4)]
        with torch.inference_mode():
            if self.variant == "full":
                stf = self.FO.StreamState(B)
                step = lambda f: self.FO.forward_spec(self.W, f, stf)
            elif self.smodel is None:
                st = self.O.StreamState(B)
                step = lambda f: self.O.forward_spec(self.W, f, st)
            else:
                caches = [torch.zeros(2, B, 16, 6, 33), torch.zeros(2, 3, B, 8, 2),
                          [[torch.zeros(B, 16, 2 * d, 33) for d in (1, 2, 4, 8)] for _ in range(2)]]   # gtcrn_micro_stream.py:618-623

                def step(f):
                    y, caches[0], caches[1], caches[2] = self.smodel(f, caches[0], caches[1], caches[2])
                    return y
            for i in range(warm):
                step(frames[i % 4])
            t0 = time.perf_counter()
            for i in range(hops):
                step(frames[i % 4])
            return (time.perf_counter() - t0) / hops


def full_variant_weights():
    """Seeded random weights of the configs[3] model as a state_dict (no checkpoint exists).  Built from the PACKAGE's own
    module (gtcrn_micro_b200.GTCRN) so that the timed arm never touches oracle/; BN statistics, PReLU slopes and LayerNorm
    affine terms are randomised too so that nothing folds away."""
    import gtcrn_micro_b200 as G
    torch.manual_seed(FULL_SEED)
    m = G.GTCRN()
    g = torch.Generator().manual_seed(FULL_SEED)
    sd = m.state_dict()
    for k, v in sd.items():
        if k.startswith("erb.") or not v.is_floating_point():
            continue
        if k.endswith("running_var"):
            sd[k] = 0.3 + 1.5 * torch.rand(v.shape, generator=g)
        elif k.endswith("running_mean") or (".bn" in k or "_bn" in k or "_ln" in k) and k.endswith("bias"):
            sd[k] = 0.3 * torch.randn(v.shape, generator=g)
        elif (".bn" in k or "_bn" in k or "_ln" in k) and k.endswith("weight"):
            sd[k] = 0.6 + 0.8 * torch.rand(v.shape, generator=g)
        elif k.endswith("act.weight"):
            sd[k] = 0.05 + 0.5 * torch.rand(v.shape, generator=g)
        elif v.dim() > 1:
            sd[k] = v * 1.7
    return {k: v.clone() for k, v in sd.items()}


def cpu_streaming(ref: "CpuReference", budget_s: float = 20.0) -> dict:
    """max real-time streams on the host CPU = B * 16 ms / t_hop, maximised over B."""
    t_start, detail, best = time.perf_counter(), {}, 0.0
    for B in (1, 64, 512, 2048):
        t = ref.stream_hop_seconds(B)
        detail[str(B)] = {"hop_ms": round(t * 1e3, 3), "realtime_streams": int(B * 0.016 / t)}
        best = max(best, B * 0.016 / t)
        if time.perf_counter() - t_start > budget_s:
            break
    return {"max_realtime_streams": int(best), "deadline_ms": 16.0, "by_batch": detail,
            "protocol": "frame-by-frame model (spectral frames in/out, zero caches), 3 warm-up + 10 timed hops per B, B*16ms/t_hop"}


def cpu_throughput(ref: "CpuReference", batches, reps: int, budget_s: float):
    """Best audio-s/s over batch sizes (SURVEY 8d protocol: 1 warm-up + best of `reps` per B)."""
    O = ref.O
    torch.set_num_threads(os.cpu_count() or 1)
    best, detail, t_start = 0.0, {}, time.perf_counter()
    for B in batches:
        x = O.synth_batch(B, L_UTT, seed=4321)
        ref.enhance(x[:1])
        times = []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            ref.enhance(x)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > budget_s:
                break
        t = min(times[1:]) if len(times) > 1 else times[0]
        detail[str(B)] = round(B * SECONDS / t, 1)
        best = max(best, B * SECONDS / t)
        if time.perf_counter() - t_start > budget_s:
            break
    return best, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.variant)
    O = ref.O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # pick the batch size that suits this host (bounded sample of the 1024-utterance workload)
    best_B, best_v = 16, 0.0
    for B in (4, 16, 64):
        x = O.synth_batch(B, L_UTT, seed=4321)
        ref.enhance(x[:1])
        t0 = time.perf_counter(); ref.enhance(x); dt = time.perf_counter() - t0
        if B * SECONDS / dt > best_v:
            best_B, best_v = B, B * SECONDS / dt
    x = O.synth_batch(best_B, L_UTT, seed=4321)
    for _ in range(args.warmup):
        ref.enhance(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.enhance(x)
    dt = time.perf_counter() - t0
    value = best_B * SECONDS * args.steps / dt
    what = "the reference's own GTCRNMicro between torch.stft / torch.istft (baseline/_ref)" if ref.kind == "reference" else "oracle port of the reference"
    if args.variant == "full":
        what = "builder-written oracle of the original-GTCRN-style variant (PARITY UNPINNED: the reference has no implementation of it)"
    sample = (f"{what}: {best_B} x {SECONDS} s utterances per step (bounded sample of the 1024-utterance batch), "
              f"torch {torch.__version__} CPU, {cores} threads")
    emit({
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * dt / args.steps, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"GTCRN-Micro offline batched inference, 1024 x {SECONDS} s synthetic 16 kHz utterances per GPU (configs[1]); CPU arm times a bounded sample"
                                if args.variant == "micro" else
                                f"original-GTCRN-style variant (configs[3], parity unpinned), 1024 x {SECONDS} s synthetic utterances per GPU; CPU arm times a bounded sample"),
                   "variant": args.variant},
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": cores, "kind": ref.kind, "sample": sample,
                         "streaming": cpu_streaming(ref)},
        "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
STREAM_HOP_BYTES = 93_000                 # minimal HBM bytes per stream-hop (2 KB audio + 88.7 KB ring-buffer state + activations)


def _hop_stats(ms):
    ms = sorted(ms)
    return {"hop_ms_p50": round(ms[len(ms) // 2], 4), "hop_ms_p99": round(ms[min(len(ms) - 1, int(0.99 * len(ms)))], 4), "hop_ms_max": round(ms[-1], 4)}


def _stream_wav_hops(smodel, N, device, hops, warm=10, groups=None):
    """`hops` real-time hops for N streams, wav in / wav out through PINNED HOST buffers: per hop H2D of [N,256] new samples,
    gtcrnb_stream_step_wav, D2H of [N,256] enhanced samples, all inside the hop's CUDA-event bracket.  The call is the package's
    StreamPool.step_wav_host: above 32,768 streams it splits them into groups whose copies overlap the other groups' kernels."""
    from gtcrn_micro_b200 import StreamPool
    pool = StreamPool(smodel, N, device, groups=groups)
    g = torch.Generator().manual_seed(N)
    hin = [torch.empty(N, 256, pin_memory=True).copy_(0.1 * torch.randn(N, 256, generator=g)) for _ in range(4)]
    hout = torch.empty(N, 256, pin_memory=True)

    def hop(i):
        pool.step_wav_host(hin[i % 4], hout)
    for i in range(warm):
        hop(i)
    torch.cuda.synchronize(device)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(hops)]
    for i in range(hops):
        evs[i][0].record()
        hop(i)
        evs[i][1].record()
    torch.cuda.synchronize(device)
    ok = bool(torch.isfinite(hout).all())
    _stream_wav_hops.last_groups = len(pool.sizes)
    del pool
    return [a.elapsed_time(b) for a, b in evs], ok


def _stream_spec_hops(smodel, N, device, hops, warm=10):
    """device-resident view: spectral frames in / out, no copies (kernel time of one hop)."""
    st = smodel.new_state(N, device)
    g = torch.Generator(device=device).manual_seed(99)
    frames = [torch.randn(N, 257, 2, device=device, generator=g) for _ in range(4)]
    out = torch.empty(N, 257, 2, device=device)
    for i in range(warm):
        smodel.step(frames[i % 4], st, out)
    torch.cuda.synchronize(device)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(hops)]
    for i in range(hops):
        evs[i][0].record()
        smodel.step(frames[i % 4], st, out)
        evs[i][1].record()
    torch.cuda.synchronize(device)
    del st
    return [a.elapsed_time(b) for a, b in evs]


def bench_streaming(G, model, device, hbm_peak, n_list=(4096, 16384), hops=625, max_search=True):
    """BASELINE configs[2]: per-hop latency (p50 / p99 / max over >= 625 hops) of the frame-by-frame path against the 16 ms
    deadline, wav-in / wav-out with the host<->device copies inside every hop, plus the device-resident kernel view and the
    largest stream count whose WORST hop still meets the deadline (doubling + bisection, confirmed over `hops` hops)."""
    res = {"deadline_ms": 16.0, "hops": hops, "io": "wav in / wav out, pinned host buffers, H2D + D2H of N x 256 samples inside every hop (StreamPool.step_wav_host)",
           "latency_contract": "output lags input by 512 samples (32 ms); gtcrnb_stream_flush_wav emits the last two hops"}
    smodel = G.StreamGTCRN() if isinstance(model, G.GTCRN) else G.StreamGTCRNMicro()
    G.convert_to_stream(smodel, model)
    per_stream_floats = 18816 if isinstance(model, G.GTCRN) else 38880
    for N in n_list:
        ms, ok = _stream_wav_hops(smodel, N, device, hops)
        r = _hop_stats(ms)
        r["realtime"] = r["hop_ms_max"] < 16.0 and ok
        spec_ms = sorted(_stream_spec_hops(smodel, N, device, 100))
        r["device_resident_spec_hop_ms_p50"] = round(spec_ms[len(spec_ms) // 2], 4)
        r["ns_per_frame"] = round(1e6 * r["device_resident_spec_hop_ms_p50"] / N, 2)
        gbs = STREAM_HOP_BYTES * N / (r["device_resident_spec_hop_ms_p50"] * 1e-3) / 1e9
        r["roofline"] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 4),
                         "bytes_per_stream_hop": STREAM_HOP_BYTES}
        res[str(N)] = r
        torch.cuda.empty_cache()
    if max_search:
        def worst(N, h):
            ms, ok = _stream_wav_hops(smodel, N, device, h, warm=5)
            torch.cuda.empty_cache()
            return max(ms) if ok else float("inf")
        free = torch.cuda.mem_get_info(device)[0]
        cap = int(free * 0.7 // (per_stream_floats * 4 + 27 * 1024 + 4096))          # state + per-hop workspace per stream
        lo, hi = max(n_list), None
        n = lo * 2
        while hi is None and n <= cap:
            if worst(n, 40) < 16.0:
                lo, n = n, n * 2
            else:
                hi = n
        if hi is None:
            hi = min(cap, n)
        while hi - lo > max(1024, lo // 32):
            mid = (lo + hi) // 2 // 256 * 256
            if worst(mid, 40) < 16.0:
                lo = mid
            else:
                hi = mid
        ms, ok = _stream_wav_hops(smodel, lo, device, hops)
        while (max(ms) >= 16.0 or not ok) and lo > max(n_list):          # confirmation over the full hop count; back off 3 % at a time
            lo = int(lo * 0.97) // 256 * 256
            ms, ok = _stream_wav_hops(smodel, lo, device, hops)
        r = _hop_stats(ms)
        r["streams"] = lo
        r["stream_groups"] = getattr(_stream_wav_hops, "last_groups", 1)
        r["memory_cap_streams"] = cap
        res["max_realtime_streams"] = r
        torch.cuda.empty_cache()
    return res


PARITY_IDX = [0, 1, 63, 64, 127, 128, 255, 256, 511, 512, 700, 767, 768, 900, 1022, 1023]


def parity_of_timed_batch(x_sub: torch.Tensor, clean_sub: torch.Tensor, y_sub: torch.Tensor, variant: str = "micro", weights=None) -> dict:
    """CHECKER leg (the only place the B200 arm touches oracle/): the oracle's CPU restatement of the reference run on a fixed
    16-utterance subset of the batch that was just timed, compared with the timed run's own output rows."""
    import gtcrn_oracle as O
    if variant == "full":
        import gtcrn_full_oracle as FO
        ref = FO.enhance(weights, x_sub)
    else:
        W = O.load_weights_npz(WEIGHTS)
        ref = O.enhance(W, x_sub)
    n = ref.shape[1]
    err = float((y_sub - ref).abs().max())
    d = max(abs(O.si_snr(clean_sub[i, :n].numpy(), y_sub[i].numpy()) - O.si_snr(clean_sub[i, :n].numpy(), ref[i].numpy())) for i in range(ref.shape[0]))
    return {"max_abs": err, "si_snr_delta_db": round(d, 6), "n": int(ref.shape[0]), "tolerance": {"max_abs": 1e-4, "si_snr_delta_db": 0.01},
            "against": ("oracle/gtcrn_oracle.py" if variant == "micro" else "oracle/gtcrn_full_oracle.py (PARITY UNPINNED: builder-written, no reference implementation)")
                       + " on the same 16 utterances of the timed batch", "ok": bool(err <= 1e-4 and d <= 0.01)}


def run_b200(args):
    import gtcrn_micro_b200 as G
    from gtcrn_micro_b200 import build
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    full = args.variant == "full"
    if full:
        full_sd = full_variant_weights()
        model = G.GTCRN()
        model.load_state_dict(full_sd)
    else:
        model = G.GTCRNMicro()
        model.load_state_dict(G.load_state_dict_npz(WEIGHTS))

    B = args.batch
    pidx = [i for i in PARITY_IDX if i < B]
    clean_rows = [None] * len(pidx)
    x = synth_on_device(B, L_UTT, device, seed=1234 + rank * B, keep_clean=(pidx, clean_rows))      # 655 MB per rank, > L2 (126 MB)
    # warm-up (also sizes the workspace) ----------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        y = model.enhance(x)
    torch.cuda.synchronize(device)
    model.profile(True); model.profile_read(reset=True)
    clocks = ClockSampler(local)
    launches0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    clocks.start()
    e0.record()
    for _ in range(args.steps):
        y = model.enhance(x)
    e1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    launches = model.launch_count() - launches0
    prof = model.profile_read(reset=True)
    model.profile(False)
    assert bool(torch.isfinite(y).all())
    x_sub, y_sub = x[pidx].cpu(), y[pidx].cpu()            # rows of the batch that was timed / of the last timed step's output

    # e2e: pinned host buffers through the C-ABI host entry point ----------------------------------
    xh = torch.empty((B, L_UTT), dtype=torch.float32, pin_memory=True)
    xh.copy_(x)
    yh = torch.empty((B, 256 * (L_UTT // 256)), dtype=torch.float32, pin_memory=True)
    del x, y
    for _ in range(2):
        model.enhance_host(xh, yh, device=device)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.enhance_host(xh, yh, device=device)          # synchronous: returns when the last D2H landed
    torch.cuda.synchronize(device)
    e2e_s = time.perf_counter() - t0

    # the same end-to-end call with PCM16 samples on the wire (wav-file workloads; half the PCIe bytes) -- reported beside e2e
    xq = torch.empty((B, L_UTT), dtype=torch.int16, pin_memory=True)
    xq.copy_((xh * 32768.0).clamp_(-32768, 32767).to(torch.int16))
    yq = torch.empty((B, 256 * (L_UTT // 256)), dtype=torch.int16, pin_memory=True)
    for _ in range(2):
        model.enhance_host_pcm16(xq, yq, device=device)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.enhance_host_pcm16(xq, yq, device=device)
    torch.cuda.synchronize(device)
    pcm_s = time.perf_counter() - t0

    times = torch.tensor([ms_total, e2e_s * 1e3, pcm_s * 1e3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)        # plumbing only: max over ranks of the device time
    ms_total, e2e_ms, pcm_ms = float(times[0]), float(times[1]), float(times[2])
    gather_info = None
    if world > 1:
        # Host-side gather of the shards (shard.gather_to_rank0; the only cross-rank data movement of the whole path, OUTSIDE the timed
        # regions): a small job of 5 one-second utterances per rank is sharded in contiguous blocks, every rank enhances its block on
        # its own GPU, rank 0 gathers the blocks over a gloo group and compares them bitwise with its own run of the whole job.
        from gtcrn_micro_b200.shard import gather_to_rank0, shard_range
        n_items = 5 * world - 2                                   # last shard short on purpose
        job = 0.1 * torch.randn(n_items, 16000, generator=torch.Generator().manual_seed(4242))
        lo, hi = shard_range(n_items, rank, world)
        mine = model.enhance(job[lo:hi].to(device)).cpu() if hi > lo else torch.zeros(0, 256 * (16000 // 256))
        gloo = dist.new_group(backend="gloo")
        t0 = time.perf_counter()
        whole = gather_to_rank0(mine, n_items, group=gloo)
        g_ms = (time.perf_counter() - t0) * 1e3
        if rank == 0:
            ref_all = model.enhance(job.to(device)).cpu()
            gather_info = {"items": n_items, "ranks": world, "backend": "gloo (host side, outside the timed regions)",
                           "bitwise_equal_to_single_gpu_run": bool(torch.equal(whole, ref_all)), "ms": round(g_ms, 2)}
    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    audio_s = world * B * SECONDS * args.steps
    value = audio_s / (ms_total * 1e-3)
    e2e_value = audio_s / (e2e_ms * 1e-3)
    peaks = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    mac_per_frame = FULL_MAC_PER_FRAME if full else MAC_PER_FRAME
    step_kernel_ms = sum(v["ms"] for v in prof.values())
    dom = max(prof, key=lambda k: prof[k]["ms"])
    whole = {"achieved_tflops": round(2.0 * mac_per_frame * FRAMES_PER_UTT * B * args.steps / (ms_total * 1e-3) / 1e12, 2),
             "frac_fp32": round(2.0 * mac_per_frame * FRAMES_PER_UTT * B * args.steps / (ms_total * 1e-3) / 1e12 / FP32_PEAK_TFLOPS, 4),
             "mac_per_frame": mac_per_frame,
             "achieved_hbm_gbs_algorithmic": round(8.0 * L_UTT * B * args.steps / (ms_total * 1e-3) / 1e9, 1),
             "frac_hbm": round(8.0 * L_UTT * B * args.steps / (ms_total * 1e-3) / 1e9 / hbm_peak, 4),
             "hbm_peak_gbs": hbm_peak, "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}
    peak_src = ("measured FFMA micro-benchmark on this pool's B200 (tools/microbench/ffma_feed.cu, profiles/ffma_feed_r1.txt); "
                "MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only and this path is FP32-FMA bound (SURVEY 8d)")
    if full:
        # dominant kernel class of the variant, against the FP32 roof (all its contractions run on the FMA pipe) and the HBM roof
        dk = prof[dom]
        cls_mac = FULL_CLASS_MAC.get(dom, 0)
        tfl = (2.0 * cls_mac * dk["frames"]) / (dk["ms"] * 1e-3) / 1e12 if dk["ms"] > 0 else 0.0
        alg_bytes = (3 if dom == "full_dec_gtconv" else 2) * 33 * 16 * 4
        roofline = {"kernel": f"{dom} ({dk['launches'] // max(1, args.steps)} launches per step)", "bound": "fp32_fma", "achieved": round(tfl, 2), "peak": FP32_PEAK_TFLOPS,
                    "unit": "TFLOP/s", "frac": round(tfl / FP32_PEAK_TFLOPS, 4), "peak_source": peak_src, "traffic": None,
                    "algorithmic_bytes_per_launch": alg_bytes * (dk["frames"] // max(1, dk["launches"])),
                    "hbm_view": {"bound": "hbm", "achieved": round(alg_bytes * dk["frames"] / (dk["ms"] * 1e-3) / 1e9, 1) if dk["ms"] > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": round(alg_bytes * dk["frames"] / (dk["ms"] * 1e-3) / 1e9 / hbm_peak, 4) if dk["ms"] > 0 else 0.0},
                    "ms_per_launch": round(dk["ms"] / max(1, dk["launches"]), 4), "launches": dk["launches"],
                    "share_of_step": round(dk["ms"] / step_kernel_ms, 4) if step_kernel_ms else None, "dominant_class_by_time": dom,
                    "whole_path": whole, "per_class_ms": {k: round(v["ms"] / args.steps, 3) for k, v in prof.items() if v["launches"]},
                    "per_class_tflops": {k: round(2.0 * FULL_CLASS_MAC[k] * v["frames"] / (v["ms"] * 1e-3) / 1e12, 2) for k, v in prof.items() if v["launches"] and k in FULL_CLASS_MAC}}
    else:
        # roofline of the dominant kernel class, timed live with CUDA events inside the library (gtcrnb_profile_*)
        tc_mask = model.tc_mask()
        dk = prof[dom]
        ci = dict(MICRO_CLASS_INFO[dom])
        on_tc = (dom == "dec_gtconv" and tc_mask & 2) or (dom == "gtcn" and tc_mask & 1)
        if on_tc:
            ci.update(MICRO_TC_INFO[dom])
        frames_per_launch = dk["frames"] // max(1, dk["launches"])
        tfl = (2.0 * ci["mac"] * dk["frames"]) / (dk["ms"] * 1e-3) / 1e12 if dk["ms"] > 0 else 0.0
        gbs = ci["alg_bytes"] * dk["frames"] / (dk["ms"] * 1e-3) / 1e9 if dk["ms"] > 0 else 0.0
        fp32_view = {"bound": "fp32_fma", "achieved": round(tfl, 2), "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": round(tfl / FP32_PEAK_TFLOPS, 4),
                     "peak_source": peak_src}
        hbm_view = {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 4),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}
        primary = hbm_view if ci["bound"] == "hbm" else fp32_view
        roofline = {
            "kernel": f"{ci['kernel']} ({dk['launches'] // max(1, args.steps)} launches per step)",
            "bound": primary["bound"], "achieved": primary["achieved"], "peak": primary["peak"], "unit": primary["unit"], "frac": primary["frac"],
            "peak_source": primary["peak_source"],
            "why_this_bound": ci["why"],
            "traffic": ci["dram_bytes"] * frames_per_launch if ci.get("dram_bytes") else None,
            "traffic_note": ci.get("dram_note"),
            "algorithmic_bytes_per_launch": ci["alg_bytes"] * frames_per_launch,
            "algorithmic_flop_per_launch": 2 * ci["mac"] * frames_per_launch,
            "other_view": fp32_view if primary is hbm_view else hbm_view,
            "ms_per_launch": round(dk["ms"] / max(1, dk["launches"]), 4), "launches": dk["launches"],
            "share_of_step": round(dk["ms"] / step_kernel_ms, 4) if step_kernel_ms else None,
            "dominant_class_by_time": dom,
            "tensor_core_layers": {"mask": tc_mask, "gtcn_pointwise": bool(tc_mask & 1), "decoder_3x3": bool(tc_mask & 2),
                                   "front_en1": bool(tc_mask & 4), "back_de3": bool(tc_mask & 8),
                                   "operands": "fp16 hi/lo (decoder, en1, de3), 3xTF32 (GTCN): fp32-class accuracy; A operand in tensor memory (decoder, GTCN) or shared memory (en1, de3)"},
            "whole_path": whole,
            "per_class_ms": {k: round(v["ms"] / args.steps, 3) for k, v in prof.items() if v["launches"]},
            "per_class": {k: {"ms": round(v["ms"] / args.steps, 3),
                              "alg_tflops": round(2.0 * MICRO_CLASS_INFO[k]["mac"] * v["frames"] / (v["ms"] * 1e-3) / 1e12, 2),
                              "alg_hbm_gbs": round(MICRO_CLASS_INFO[k]["alg_bytes"] * v["frames"] / (v["ms"] * 1e-3) / 1e9, 1),
                              "frac_fp32": round(2.0 * MICRO_CLASS_INFO[k]["mac"] * v["frames"] / (v["ms"] * 1e-3) / 1e12 / FP32_PEAK_TFLOPS, 4),
                              "frac_hbm": round(MICRO_CLASS_INFO[k]["alg_bytes"] * v["frames"] / (v["ms"] * 1e-3) / 1e9 / hbm_peak, 4)}
                          for k, v in prof.items() if v["launches"] and v["ms"] > 0 and k in MICRO_CLASS_INFO},
        }
    wl = (f"GTCRN-Micro offline batched inference, {B} x {SECONDS} s synthetic 16 kHz utterances per GPU (BASELINE configs[1]), STFT+model+iSTFT, shipped DNS3 checkpoint weights"
          if not full else
          f"original-GTCRN-style variant (BASELINE configs[3]: unfold SFE, groups=2, GTConv dilations 1/2/5, GRU-TRA, DPGRNN; PARITY UNPINNED, seeded random weights), "
          f"{B} x {SECONDS} s synthetic 16 kHz utterances per GPU, STFT+model+iSTFT")
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "variant": args.variant,
                   "batch_per_gpu": B, "samples_per_utt": L_UTT, "frames_per_utt": FRAMES_PER_UTT,
                   "cache": f"inputs {B * L_UTT * 4 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                   "parallelism": f"{world} independent shard(s), no collective"},
        "clocks": clk,
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": B * L_UTT * 4, "d2h_bytes_per_step": B * 256 * (L_UTT // 256) * 4,
                "ms_per_step": round(e2e_ms / args.steps, 3), "api": "GTCRNMicro.enhance_host -> gtcrnb_enhance_host (pinned host in/out)"},
        "e2e_pcm16": {"value": round(audio_s / (pcm_ms * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": B * L_UTT * 2,
                      "d2h_bytes_per_step": B * 256 * (L_UTT // 256) * 2, "ms_per_step": round(pcm_ms / args.steps, 3),
                      "api": "GTCRNMicro.enhance_host_pcm16 -> gtcrnb_enhance_host_pcm16 (pinned int16 host in/out, decode/encode on device)"},
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if gather_info is not None:
        line["host_gather"] = gather_info
    if not args.no_stream:
        try:
            line["streaming"] = bench_streaming(G, model, device, hbm_peak, max_search=(world == 1))
        except Exception as e:  # streaming is reported, never allowed to sink the offline number
            line["streaming"] = {"error": f"{type(e).__name__}: {e}"}
    if not args.no_parity:
        line["parity"] = parity_of_timed_batch(x_sub, torch.stack(clean_rows), y_sub, args.variant, full_sd if full else None)
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ref = CpuReference(args.variant, full_sd if full else None)
        v, detail = cpu_throughput(ref, (1, 4, 16, 64), reps=2, budget_s=45.0)
        what = "the reference's own GTCRNMicro between torch.stft / torch.istft (baseline/_ref)" if ref.kind == "reference" else "oracle port (same ATen ops as the reference)"
        if full:
            what = "builder-written oracle of the original-GTCRN-style variant (PARITY UNPINNED: no reference implementation exists)"
        line["cpu_baseline"] = {"value": round(v, 1), "unit": UNIT, "cores": cores, "kind": ref.kind,
                                "sample": f"{what}, torch {torch.__version__} CPU, {SECONDS} s utterances, "
                                          f"best of B in {detail} (1 warm-up + best of 2 each); 1024-utterance batch extrapolated from this sample",
                                "streaming": cpu_streaming(ref)}
        if "streaming" in line and "max_realtime_streams" in line["streaming"]:
            line["streaming"]["vs_host_cpu"] = round(line["streaming"]["max_realtime_streams"]["streams"] / max(1, line["cpu_baseline"]["streaming"]["max_realtime_streams"]), 1)
    emit(line)
    if "parity" in line and not line["parity"]["ok"]:
        sys.stderr.write(f"bench: PARITY FAILURE on the timed batch: {line['parity']}\n")
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        sys.exit(3)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(data); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="utterances per GPU per step (BASELINE configs[1]: 1024)")
    ap.add_argument("--no-stream", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--variant", default="micro", choices=["micro", "full"],
                    help="micro: the reference's GTCRNMicro (BASELINE configs[1], the headline); full: the original-GTCRN-style architecture of configs[3] (parity unpinned)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed batch (A/B runs only)")
    ap.add_argument("--sweep-utterances", type=int, default=0,
                    help="BASELINE configs[4]: a fixed job of this many 10 s utterances (65536) sharded over the ranks in contiguous "
                         "blocks (shard.shard_range); sets --steps to the rank's share / --batch and reports scaling 'strong'")
    args = ap.parse_args()
    args.scaling = "weak"
    if args.sweep_utterances:
        from gtcrn_micro_b200.shard import shard_range
        world = int(os.environ.get("WORLD_SIZE", "1"))
        lo, hi = shard_range(args.sweep_utterances, 0, world)          # rank 0 holds the largest shard
        args.steps = max(1, -(-(hi - lo) // args.batch))
        args.scaling = "strong"
    # stdout carries exactly ONE JSON line: park the real stdout and point fd 1 at stderr while the libraries run
    # (NCCL prints its version banner on stdout), then hand the real stdout back to emit().
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
