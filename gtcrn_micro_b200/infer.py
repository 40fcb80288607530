"""Batched file pipeline: the B200 replacement for the reference's per-file inference loop.

Mirrors /root/reference/gtcrn_micro/infer.py:26-119: same inputs (a folder of noisy wavs, a folder of
clean references matched by DNS3 `fileid_N`), same outputs (`<uid>_enh.wav` PCM16 files, `inf.scp`, `ref.scp`), same
length fix-up against the clean reference (pad / truncate, infer.py:97-102), same YAML keys when driven by `-C`.
What changes is the schedule: files are bucketed by exact length (no padding, so results are bit-identical to running each
file alone); a bucket of PCM16 files goes through `enhance_host_pcm16` (C ABI gtcrnb_enhance_host_pcm16): PCM16 over PCIe in
both directions, decoded, enhanced by the fused wav->wav kernels and quantised on the device (floor(y * 32768), the rule the
reference's shipped *_enh.wav files follow), then length-fixed and written out.  Optionally scores the pairs on the device (eval/eval_intrusive_metrics.py SI-SNR / SDR).

    python -m gtcrn_micro_b200.infer --noisy-dir N --clean-dir C --enh-dir E --checkpoint ckpt.tar [--metrics]
    python -m gtcrn_micro_b200.infer -C infer.yaml            # reference layout: test_dataset.{noisy_dir,clean_dir}, network.{enh_folder,checkpoint}

Resampling (the reference calls librosa for fs != 16 kHz) is not part of the hot path and is refused with an error.
"""
from __future__ import annotations

import argparse
import os
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

FS = 16000


@dataclass
class Job:
    uid: str
    noisy_path: str
    ref_path: Optional[str]


def extract_fileid(path: str) -> Optional[str]:
    """infer.py:16-22"""
    base = os.path.basename(path)
    if "fileid_" not in base:
        return None
    return base.split("fileid_")[-1].split(".")[0]


def list_jobs(noisy_dir: str, clean_dir: Optional[str]) -> List[Job]:
    """Sorted *.wav of noisy_dir, each matched to clean_fileid_<id>.wav (infer.py:41-90).  clean_dir=None skips matching."""
    jobs = []
    for name in sorted(n for n in os.listdir(noisy_dir) if n.endswith("wav")):
        noisy_path = os.path.join(noisy_dir, name)
        ref_path = None
        if clean_dir is not None:
            fid = extract_fileid(noisy_path)
            if fid is None:
                raise RuntimeError(f"Unable to extract: {noisy_path}")
            ref_path = os.path.join(clean_dir, f"clean_fileid_{fid}.wav")
            if not os.path.exists(ref_path):
                raise FileNotFoundError(f"Clean file not found for clean_fileid_{fid}.wav, fileid={fid}:\n {ref_path}")
        jobs.append(Job(name.split(".wav")[0], noisy_path, ref_path))
    return jobs


def read_wav(path: str) -> np.ndarray:
    """Mono 16 kHz wav -> int16 (PCM16 files, kept as PCM for the H2D copy) or float32 samples."""
    from scipy.io import wavfile
    fs, data = wavfile.read(path)
    if fs != FS:
        raise ValueError(f"{path}: sample rate {fs} != {FS}; resample first (not part of the B200 path)")
    if data.ndim != 1:
        raise ValueError(f"{path}: expected mono audio, got shape {data.shape}")
    if data.dtype == np.int16:
        return data
    if data.dtype == np.int32:
        return (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    if data.dtype == np.uint8:
        return ((data.astype(np.float32) - 128.0) / 128.0).astype(np.float32)
    return data.astype(np.float32)


def wav_num_samples(path: str) -> int:
    """Length from the RIFF header only (no sample data read): used to bucket before loading."""
    import struct
    with open(path, "rb") as f:
        head = f.read(12)
        if head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file")
        block_align = None
        while True:
            ck = f.read(8)
            if len(ck) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = ck[:4], struct.unpack("<I", ck[4:])[0]
            if cid == b"fmt ":
                fmt = f.read(size + (size & 1))
                block_align = struct.unpack("<H", fmt[12:14])[0]
            elif cid == b"data":
                if not block_align:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                return size // block_align
            else:
                f.seek(size + (size & 1), 1)


def bucket_by_length(lengths: Sequence[int], max_batch: int) -> List[List[int]]:
    """Indices grouped by identical length, each group split into batches of <= max_batch; order inside a group is input order."""
    groups: Dict[int, List[int]] = {}
    for i, n in enumerate(lengths):
        groups.setdefault(int(n), []).append(i)
    out = []
    for n in sorted(groups):
        idx = groups[n]
        out += [idx[i:i + max_batch] for i in range(0, len(idx), max_batch)]
    return out


def fix_length(enh: torch.Tensor, n: int) -> torch.Tensor:
    """[B,L] -> [B,n]: zero-pad or truncate to the clean reference's length (infer.py:97-102)."""
    L = enh.shape[1]
    if L < n:
        return torch.nn.functional.pad(enh, (0, n - L))
    return enh[:, :n]


def to_pcm16(x: torch.Tensor) -> torch.Tensor:
    return torch.floor(x * 32768.0).clamp_(-32768.0, 32767.0).to(torch.int16)


def write_scp(path: str, rows: Sequence[Tuple[str, str]]) -> None:
    with open(path, "w") as f:
        for uid, p in rows:
            f.write(f"{uid} {p}\n")


@torch.no_grad()
def run(model, jobs: Sequence[Job], enh_dir: str, max_batch: int = 256, device: str = "cuda:0", metrics: bool = False):
    """Enhance every job; returns (inf_scp_rows, ref_scp_rows, scores) with scores = {uid: (si_snr, sdr)} when metrics=True."""
    from scipy.io import wavfile
    os.makedirs(enh_dir, exist_ok=True)
    dev = torch.device(device)
    lengths = [wav_num_samples(j.noisy_path) for j in jobs]
    enh_paths: List[Optional[str]] = [None] * len(jobs)
    scores: Dict[str, Tuple[float, float]] = {}
    def to_f32(a: np.ndarray) -> np.ndarray:
        return a.astype(np.float32) / 32768.0 if a.dtype == np.int16 else a

    def fix_np(y: np.ndarray, n: Optional[int]) -> np.ndarray:           # infer.py:97-102 on one PCM16 row
        if n is None or n == len(y):
            return y
        return np.pad(y, (0, n - len(y))) if len(y) < n else y[:n]

    for batch in bucket_by_length(lengths, max_batch):
        raw = [read_wav(jobs[i].noisy_path) for i in batch]
        refs = [read_wav(jobs[i].ref_path) if jobs[i].ref_path else None for i in batch]
        ref_n = [None if r is None else len(r) for r in refs]
        if all(r.dtype == np.int16 for r in raw) and dev.type == "cuda" and hasattr(model, "enhance_host_pcm16"):
            # PCM16 end to end: 2 bytes per sample each way over PCIe, decode / enhance / quantise on the device
            pcm_in = torch.from_numpy(np.stack(raw)).pin_memory()
            pcm = model.enhance_host_pcm16(pcm_in, device=dev).numpy()
        else:
            x = torch.from_numpy(np.stack([to_f32(r) for r in raw])).to(dev)
            y = model.enhance(x)
            pcm = (model.float_to_pcm16(y) if hasattr(model, "float_to_pcm16") and dev.type == "cuda" else to_pcm16(y)).cpu().numpy()
        outs = [fix_np(pcm[k], ref_n[k]) for k in range(len(batch))]
        if metrics:
            # scored on what was written (the evaluation script reads the PCM16 files back): group rows of equal length
            by_len: Dict[int, List[int]] = {}
            for k in range(len(batch)):
                if refs[k] is not None:
                    by_len.setdefault(len(outs[k]), []).append(k)
            for ks in by_len.values():
                r = torch.from_numpy(np.stack([to_f32(refs[k]) for k in ks])).to(dev)
                e16 = torch.from_numpy(np.stack([outs[k] for k in ks])).to(dev)
                e = model.pcm16_to_float(e16) if hasattr(model, "pcm16_to_float") and dev.type == "cuda" else e16.to(torch.float32).mul_(1.0 / 32768.0)
                sc = model.pair_metrics(r, e).cpu().numpy()
                for row, k in enumerate(ks):
                    scores[jobs[batch[k]].uid] = (float(sc[row, 0]), float(sc[row, 1]))
        for k, i in enumerate(batch):
            enh_paths[i] = os.path.join(enh_dir, jobs[i].uid + "_enh.wav")
            wavfile.write(enh_paths[i], FS, outs[k])
    inf_rows = [(j.uid, enh_paths[i]) for i, j in enumerate(jobs)]
    ref_rows = [(j.uid, j.ref_path) for j in jobs if j.ref_path]
    write_scp(os.path.join(enh_dir, "inf.scp"), inf_rows)
    if ref_rows:
        write_scp(os.path.join(enh_dir, "ref.scp"), ref_rows)
    return inf_rows, ref_rows, scores


_INTERP = re.compile(r"\$\{([^${}]+)\}")


def load_config(path: str) -> dict:
    """Reference-style YAML with OmegaConf `${a.b.c}` interpolations resolved (the reference reads cfg_infer.yaml with
    OmegaConf.load, infer.py:27; its `checkpoint` / `enh_folder` entries interpolate `network.exp_path` / `network.ckpt_name`).
    Absolute dotted keys only (what the reference's configs use); nested references resolve recursively; cycles raise."""
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}

    def lookup(key: str):
        node = cfg
        for part in key.strip().split("."):
            if not isinstance(node, dict) or part not in node:
                raise KeyError(f"{path}: interpolation key '{key}' not found")
            node = node[part]
        return node

    def resolve(v, stack=()):
        if isinstance(v, dict):
            return {k: resolve(x, stack) for k, x in v.items()}
        if isinstance(v, list):
            return [resolve(x, stack) for x in v]
        if not isinstance(v, str) or "${" not in v:
            return v
        whole = _INTERP.fullmatch(v)
        if whole:                                                   # "${a.b}" alone keeps the referenced value's type
            key = whole.group(1)
            if key in stack:
                raise ValueError(f"{path}: interpolation cycle through '{key}'")
            return resolve(lookup(key), stack + (key,))

        def sub(m):
            key = m.group(1)
            if key in stack:
                raise ValueError(f"{path}: interpolation cycle through '{key}'")
            return str(resolve(lookup(key), stack + (key,)))
        return _INTERP.sub(sub, v)

    return resolve(cfg)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-C", "--config", default=None, help="reference-style infer YAML")
    ap.add_argument("-D", "--device", default="0")
    ap.add_argument("--noisy-dir"); ap.add_argument("--clean-dir"); ap.add_argument("--enh-dir"); ap.add_argument("--checkpoint")
    ap.add_argument("--max-batch", type=int, default=256)
    ap.add_argument("--metrics", action="store_true", help="score SI-SNR / SDR on the device against the clean references")
    args = ap.parse_args(argv)
    if args.config:
        cfg = load_config(args.config)
        args.noisy_dir = args.noisy_dir or cfg["test_dataset"]["noisy_dir"]
        args.clean_dir = args.clean_dir or cfg["test_dataset"].get("clean_dir")
        args.enh_dir = args.enh_dir or cfg["network"]["enh_folder"]
        args.checkpoint = args.checkpoint or cfg["network"]["checkpoint"]
    if not (args.noisy_dir and args.enh_dir and args.checkpoint):
        ap.error("need --noisy-dir, --enh-dir and --checkpoint (or -C)")
    from .modules import load_checkpoint
    model = load_checkpoint(args.checkpoint)
    jobs = list_jobs(args.noisy_dir, args.clean_dir)
    _, _, scores = run(model, jobs, args.enh_dir, args.max_batch, f"cuda:{args.device}", args.metrics)
    print(f"enhanced {len(jobs)} files -> {args.enh_dir}")
    if scores:
        s = np.array(list(scores.values()))
        print(f"SI-SNR: {s[:, 0].mean():.4f} dB   SDR: {s[:, 1].mean():.4f} dB   ({len(scores)} pairs)")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
