"""nn.Module surface of the B200 path: same constructors, `forward` contracts and `state_dict` keys as
the reference, every computation delegated to libgtcrn_b200.so (hand-written sm_100a kernels).

  GTCRNMicro              gtcrn_micro/models/gtcrn_micro.py:485-532
  StreamGTCRNMicro        gtcrn_micro/streaming/gtcrn_micro_stream.py:521-574
  convert_to_stream       gtcrn_micro/streaming/conversion/convert.py:7-56
  enhance / enhance_host  the STFT -> model -> iSTFT core of gtcrn_micro/infer.py:60-76 (batched)

The modules hold parameters/buffers only (no nn.Conv2d etc.): a checkpoint is loaded with the usual
`model.load_state_dict(torch.load(path)["model"])`, packed on the host (pack.py) and handed to the
library.  Inference only (no autograd).  No CPU or PyTorch fallback exists: tensors must live on a
CUDA device of compute capability 10.x and the shared library must have been built.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import lib as _lib
from . import pack as _pack
from .erb import erb_filter_bank

_BN = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")
TCN_DILATIONS = (1, 2, 4, 8)


# ------------------------------------------------------------------------------------------------
# parameter tree: (key, shape, kind)   kind: "w" conv-like weight, "b" bias, "bn_*", "prelu", "erb", "erbT"
# ------------------------------------------------------------------------------------------------
def _bn(prefix: str, n: int):
    return [(f"{prefix}.weight", (n,), "bn_w"), (f"{prefix}.bias", (n,), "bn_b"), (f"{prefix}.running_mean", (n,), "bn_m"),
            (f"{prefix}.running_var", (n,), "bn_v"), (f"{prefix}.num_batches_tracked", (), "bn_n")]


def _conv_block(p: str, cin: int, cout: int, deconv: bool, last: bool):
    wshape = (cin, cout, 1, 5) if deconv else (cout, cin, 1, 5)
    out = [(f"{p}.conv.weight", wshape, "w"), (f"{p}.conv.bias", (cout,), "b")] + _bn(f"{p}.bn", cout)
    if not last:
        out.append((f"{p}.act.weight", (1,), "prelu"))
    return out


def _gt_block(p: str, deconv: bool, stream: bool):
    infix = (".ConvTranspose2d" if deconv else ".Conv2d") if stream else ""
    # the streaming twin stores the decoder 3x3 as a flipped Conv2d weight [out,in,kt,kf]; same shape (16,16,3,3)
    out = [(f"{p}.point_conv1.weight", (8, 16, 1, 1) if deconv else (16, 8, 1, 1), "w"), (f"{p}.point_conv1.bias", (16,), "b")]
    out += _bn(f"{p}.point_bn1", 16) + [(f"{p}.point_act.weight", (1,), "prelu")]
    out += [(f"{p}.depth_conv{infix}.weight", (16, 16, 3, 3) if deconv else (16, 1, 3, 3), "w"), (f"{p}.depth_conv{infix}.bias", (16,), "b")]
    out += _bn(f"{p}.depth_bn", 16) + [(f"{p}.depth_act.weight", (1,), "prelu")]
    out += [(f"{p}.point_conv2.weight", (16, 8, 1, 1) if deconv else (8, 16, 1, 1), "w"), (f"{p}.point_conv2.bias", (8,), "b")]
    out += _bn(f"{p}.point_bn2", 8)
    out += [(f"{p}.tra.depth_conv.weight", (8, 1, 3), "w"), (f"{p}.tra.depth_conv.bias", (8,), "b"),
            (f"{p}.tra.point_conv.weight", (8, 8, 1), "w"), (f"{p}.tra.point_conv.bias", (8,), "b")]
    return out


def _tcn(p: str, stream: bool):
    infix = ".Conv2d" if stream else ""
    out = [(f"{p}.conv1.weight", (16, 16, 1, 1), "w"), (f"{p}.conv1.bias", (16,), "b")] + _bn(f"{p}.bn1", 16) + [(f"{p}.act1.weight", (1,), "prelu")]
    out += [(f"{p}.conv2{infix}.weight", (16, 1, 3, 1), "w"), (f"{p}.conv2{infix}.bias", (16,), "b")] + _bn(f"{p}.bn2", 16) + [(f"{p}.act2.weight", (1,), "prelu")]
    out += [(f"{p}.conv3.weight", (16, 16, 1, 1), "w"), (f"{p}.conv3.bias", (16,), "b")] + _bn(f"{p}.bn3", 16) + [(f"{p}.act3.weight", (1,), "prelu")]
    return out


def param_spec(stream: bool):
    spec = [("erb.erb_fc.weight", (64, 192), "erb"), ("erb.ierb_fc.weight", (192, 64), "erbT"),
            ("sfe.depth_conv.weight", (3, 1, 1, 3), "w")]
    spec += _conv_block("encoder.en_convs.0", 3, 16, False, False) + _conv_block("encoder.en_convs.1", 16, 16, False, False)
    for i in (2, 3, 4):
        spec += _gt_block(f"encoder.en_convs.{i}", False, stream)
    for g in (1, 2):
        for k in range(4):
            spec += _tcn(f"gtcn{g}.blocks.{k}", stream)
    for i in (0, 1, 2):
        spec += _gt_block(f"decoder.de_convs.{i}", True, stream)
    spec += _conv_block("decoder.de_convs.3", 16, 16, True, False) + _conv_block("decoder.de_convs.4", 16, 2, True, True)
    return spec


class _Node(nn.Module):
    """Pure container: gives parameters their dotted reference names."""


def _attach(root: nn.Module, key: str, tensor: torch.Tensor, kind: str) -> None:
    parts = key.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    if kind.startswith("bn_m") or kind.startswith("bn_v") or kind == "bn_n":
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def _init_tensor(shape, kind: str, gen: torch.Generator) -> torch.Tensor:
    if kind == "erb":
        return torch.from_numpy(erb_filter_bank())
    if kind == "erbT":
        return torch.from_numpy(np.ascontiguousarray(erb_filter_bank().T))
    if kind in ("bn_w", "bn_v", "ln_w"):
        return torch.ones(shape)
    if kind in ("bn_b", "bn_m", "ln_b"):
        return torch.zeros(shape)
    if kind == "bn_n":
        return torch.zeros((), dtype=torch.long)
    if kind == "prelu":
        return torch.full(shape, 0.25)
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
    bound = 1.0 / np.sqrt(max(1, fan_in))
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


# ------------------------------------------------------------------------------------------------
class _B200Module(nn.Module):
    _stream = False
    _variant = 0                  # C-ABI model variant (GTCRNB_VARIANT_MICRO / _FULL)

    @classmethod
    def _param_spec(cls):
        return param_spec(cls._stream)

    def _pack(self) -> np.ndarray:
        return _pack.pack_state_dict(self.state_dict())

    def __init__(self, n_fft: int = 512, hop_len: int = 256, win_len: int = 512):
        # like the reference constructor (gtcrn_micro.py:486-504) the three arguments are accepted for
        # config compatibility; the architecture is fixed at n_fft 512 / hop 256.
        super().__init__()
        gen = torch.Generator().manual_seed(torch.initial_seed() % (2 ** 31))
        for key, shape, kind in self._param_spec():
            _attach(self, key, _init_tensor(shape, kind, gen), kind)
        self._handle = C.c_void_p(None)
        self._handle_key = None
        self.eval()

    # -- native handle management --------------------------------------------------------------
    def _weights_key(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def _ensure_handle(self, device: torch.device) -> C.c_void_p:
        if device.type != "cuda":
            raise _lib.GtcrnB200Error("gtcrn_micro_b200 has no CPU path: tensors must be on a CUDA (sm_100a) device")
        key = (device.index if device.index is not None else torch.cuda.current_device(), self._weights_key())
        if self._handle.value is not None and key == self._handle_key:
            return self._handle
        L = _lib.load()
        self._release()
        blob = self._pack()
        with torch.cuda.device(device):
            h = C.c_void_p(None)
            _lib.check(L.gtcrnb_create_variant(blob.ctypes.data_as(C.c_void_p), blob.size, self._variant, C.byref(h)), "gtcrnb_create_variant")
        self._handle, self._handle_key = h, key
        return h

    def _release(self):
        if getattr(self, "_handle", None) is not None and self._handle.value is not None:
            try:
                _lib.load().gtcrnb_destroy(self._handle)
            except Exception:
                pass
            self._handle = C.c_void_p(None)
            self._handle_key = None

    def __del__(self):
        try:
            self._release()
        except Exception:       # interpreter shutdown: torch's module machinery may already be torn down
            pass

    def profile(self, on: bool) -> None:
        """Bracket every kernel launch of this model with CUDA events (see `profile_read`)."""
        if self._handle.value is None:
            raise _lib.GtcrnB200Error("run the model once before enabling profiling")
        _lib.check(_lib.load().gtcrnb_profile_enable(self._handle, int(on)), "gtcrnb_profile_enable")

    def profile_read(self, reset: bool = True) -> dict:
        """{kernel class: {"ms": total device ms, "launches": n, "frames": frames processed}}"""
        L = _lib.load()
        n = L.gtcrnb_profile_num_classes()
        ms, ln, fr = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_int64 * n)()
        _lib.check(L.gtcrnb_profile_read(self._handle, int(reset), ms, ln, fr), "gtcrnb_profile_read")
        return {L.gtcrnb_profile_class_name(i).decode(): {"ms": ms[i], "launches": ln[i], "frames": fr[i]} for i in range(n)}

    def launch_count(self) -> int:
        return int(_lib.load().gtcrnb_launch_count(self._handle)) if self._handle.value is not None else 0

    def tc_mask(self) -> int:
        """Layers running on the tensor cores: bit 0 GTCN pointwise convs, bit 1 decoder GTConv 3x3, bit 2 en1, bit 3 de3 (GTCRNB_TC, default 15)."""
        return int(_lib.load().gtcrnb_tc_mask(self._handle)) if self._handle.value is not None else 0

    @staticmethod
    def _check_spec(spec: torch.Tensor, T: Optional[int] = None):
        if spec.dim() != 4 or spec.shape[1] != 257 or spec.shape[3] != 2:
            raise ValueError(f"spec must be [B,257,T,2], got {tuple(spec.shape)}")
        if spec.dtype != torch.float32:
            raise TypeError("spec must be float32")
        if T is not None and spec.shape[2] != T:
            raise ValueError(f"streaming forward takes T = {T} frame per call, got {spec.shape[2]}")


class GTCRNMicro(_B200Module):
    """Drop-in for the reference `GTCRNMicro` (offline, whole utterance)."""

    @torch.no_grad()
    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        """spec [B,257,T,2] float32 CUDA (any strides, e.g. torch.stft output) -> enhanced spec [B,257,T,2]
        (a permuted view of a [B,T,257,2] buffer, like the reference returns a permuted view)."""
        self._check_spec(spec)
        h = self._ensure_handle(spec.device)
        B, _, T, _ = spec.shape
        out = torch.empty((B, T, 257, 2), dtype=torch.float32, device=spec.device).permute(0, 2, 1, 3)
        L = _lib.load()
        with torch.cuda.device(spec.device):
            _lib.check(L.gtcrnb_forward_spec(h, spec.data_ptr(), _lib.strides4(spec.stride()), B, T, out.data_ptr(),
                                             _lib.strides4(out.stride()), torch.cuda.current_stream().cuda_stream),
                       "gtcrnb_forward_spec")
        return out

    @torch.no_grad()
    def enhance(self, wav: torch.Tensor) -> torch.Tensor:
        """wav [B,L] (or [L]) float32 CUDA -> enhanced wav [B, 256*(L//256)]: fused STFT -> model -> iSTFT."""
        squeeze = wav.dim() == 1
        if squeeze:
            wav = wav[None]
        if wav.dim() != 2 or wav.dtype != torch.float32:
            raise ValueError("wav must be float32 [B,L]")
        if wav.stride(1) != 1:
            wav = wav.contiguous()
        h = self._ensure_handle(wav.device)
        B, Ls = wav.shape
        out = torch.empty((B, 256 * (Ls // 256)), dtype=torch.float32, device=wav.device)
        L = _lib.load()
        with torch.cuda.device(wav.device):
            _lib.check(L.gtcrnb_enhance(h, wav.data_ptr(), B, Ls, wav.stride(0), out.data_ptr(), out.stride(0),
                                        torch.cuda.current_stream().cuda_stream), "gtcrnb_enhance")
        return out[0] if squeeze else out

    @torch.no_grad()
    def enhance_host(self, wav: torch.Tensor, out: Optional[torch.Tensor] = None, device: Optional[torch.device] = None) -> torch.Tensor:
        """Same as `enhance` for HOST tensors (ideally pinned): chunked H2D / compute / D2H pipeline inside the library."""
        if wav.is_cuda or wav.dim() != 2 or wav.dtype != torch.float32 or not wav.is_contiguous():
            raise ValueError("wav must be a contiguous float32 CPU tensor [B,L]")
        B, Ls = wav.shape
        if out is not None and (out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (B, 256 * (Ls // 256))):
            raise ValueError("out must be a contiguous float32 CPU tensor [B,256*(L//256)]")   # the library writes B*256*(L//256) floats through this pointer
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        h = self._ensure_handle(dev)
        if out is None:
            out = torch.empty((B, 256 * (Ls // 256)), dtype=torch.float32, pin_memory=True)
        L = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(L.gtcrnb_enhance_host(h, wav.data_ptr(), B, Ls, out.data_ptr()), "gtcrnb_enhance_host")
        return out

    @torch.no_grad()
    def enhance_host_pcm16(self, wav: torch.Tensor, out: Optional[torch.Tensor] = None, device: Optional[torch.device] = None) -> torch.Tensor:
        """`enhance_host` for PCM16 audio: wav int16 [B,L] on the host (ideally pinned) -> int16 [B,256*(L//256)].
        Samples cross PCIe as 2 bytes; decode (x/32768) and encode (floor(y*32768), clamped) run on the device."""
        if wav.is_cuda or wav.dim() != 2 or wav.dtype != torch.int16 or not wav.is_contiguous():
            raise ValueError("wav must be a contiguous int16 CPU tensor [B,L]")
        B, Ls = wav.shape
        if out is not None and (out.is_cuda or out.dtype != torch.int16 or not out.is_contiguous() or tuple(out.shape) != (B, 256 * (Ls // 256))):
            raise ValueError("out must be a contiguous int16 CPU tensor [B,256*(L//256)]")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        h = self._ensure_handle(dev)
        if out is None:
            out = torch.empty((B, 256 * (Ls // 256)), dtype=torch.int16, pin_memory=True)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().gtcrnb_enhance_host_pcm16(h, wav.data_ptr(), B, Ls, out.data_ptr()), "gtcrnb_enhance_host_pcm16")
        return out

    @torch.no_grad()
    def pcm16_to_float(self, pcm: torch.Tensor) -> torch.Tensor:
        """int16 CUDA tensor -> float32 (x / 32768, what sf.read(dtype="float32") gives, infer.py:52)."""
        if not pcm.is_cuda or pcm.dtype != torch.int16:
            raise ValueError("pcm must be an int16 CUDA tensor")
        pcm = pcm.contiguous()
        out = torch.empty(pcm.shape, dtype=torch.float32, device=pcm.device)
        h = self._ensure_handle(pcm.device)
        with torch.cuda.device(pcm.device):
            _lib.check(_lib.load().gtcrnb_pcm16_to_f32(h, pcm.data_ptr(), out.data_ptr(), pcm.numel(), torch.cuda.current_stream().cuda_stream), "gtcrnb_pcm16_to_f32")
        return out

    @torch.no_grad()
    def float_to_pcm16(self, x: torch.Tensor) -> torch.Tensor:
        """float32 CUDA tensor -> int16 with the quantisation of the reference's written wav files: floor(x * 32768), clamped."""
        if not x.is_cuda or x.dtype != torch.float32:
            raise ValueError("x must be a float32 CUDA tensor")
        x = x.contiguous()
        out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
        h = self._ensure_handle(x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().gtcrnb_f32_to_pcm16(h, x.data_ptr(), out.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream), "gtcrnb_f32_to_pcm16")
        return out

    @torch.no_grad()
    def pair_metrics(self, ref: torch.Tensor, inf: torch.Tensor) -> torch.Tensor:
        """SI-SNR and SDR (dB) of waveform pairs on the device, as eval_intrusive_metrics.py:74-94 defines them.
        ref, inf [B,L] float32 CUDA -> [B,2] = (si_snr, sdr)."""
        if ref.shape != inf.shape or ref.dim() != 2 or ref.dtype != torch.float32 or inf.dtype != torch.float32:
            raise ValueError("ref and inf must be float32 [B,L] of the same shape")
        ref = ref if ref.stride(1) == 1 else ref.contiguous()
        inf = inf if inf.stride(1) == 1 else inf.contiguous()
        h = self._ensure_handle(ref.device)
        out = torch.empty((ref.shape[0], 2), dtype=torch.float32, device=ref.device)
        with torch.cuda.device(ref.device):
            _lib.check(_lib.load().gtcrnb_pair_metrics(h, ref.data_ptr(), ref.stride(0), inf.data_ptr(), inf.stride(0), ref.shape[0],
                                                       ref.shape[1], out.data_ptr(), torch.cuda.current_stream().cuda_stream), "gtcrnb_pair_metrics")
        return out

    @torch.no_grad()
    def stft(self, wav: torch.Tensor) -> torch.Tensor:
        """torch.stft(wav,512,256,512,sqrt-hann,return_complex=False) replacement: [B,L] -> [B,257,T,2] (view)."""
        h = self._ensure_handle(wav.device)
        wav = wav if wav.stride(1) == 1 else wav.contiguous()
        B, Ls = wav.shape
        T = 1 + Ls // 256
        spec = torch.empty((B, T, 257, 2), dtype=torch.float32, device=wav.device)
        with torch.cuda.device(wav.device):
            _lib.check(_lib.load().gtcrnb_stft(h, wav.data_ptr(), B, Ls, wav.stride(0), spec.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "gtcrnb_stft")
        return spec.permute(0, 2, 1, 3)

    @torch.no_grad()
    def istft(self, spec: torch.Tensor) -> torch.Tensor:
        """torch.istft replacement: spec [B,257,T,2] -> wav [B, 256*(T-1)]."""
        self._check_spec(spec)
        h = self._ensure_handle(spec.device)
        s = spec.permute(0, 2, 1, 3).contiguous()      # [B,T,257,2]
        B, T = s.shape[0], s.shape[1]
        wav = torch.empty((B, 256 * (T - 1)), dtype=torch.float32, device=spec.device)
        with torch.cuda.device(spec.device):
            _lib.check(_lib.load().gtcrnb_istft(h, s.data_ptr(), B, T, wav.data_ptr(), wav.stride(0),
                                                torch.cuda.current_stream().cuda_stream), "gtcrnb_istft")
        return wav

    @torch.no_grad()
    def debug_activation(self, name: str, B: int, T: int) -> torch.Tensor:
        """Intermediate activation of the last forward/enhance call, [B,T,F,16] channel-last (tests only)."""
        Fq = 65 if name == "en0" else 33
        dev = torch.device("cuda", self._handle_key[0])
        out = torch.empty((B, T, Fq, 16), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().gtcrnb_debug_activation(self._handle, name.encode(), out.data_ptr(), out.numel(),
                                                           torch.cuda.current_stream().cuda_stream), "gtcrnb_debug_activation")
        return out


class StreamState:
    """Opaque per-batch streaming state living in one device buffer (conv / TRA / TCN ring buffers of N
    streams plus the STFT / overlap-add tails used by `step_wav`)."""

    def __init__(self, n_streams: int, device: torch.device, variant: int = 0):
        self.n = int(n_streams)
        self.variant = int(variant)
        nbytes = int(_lib.load().gtcrnb_stream_state_bytes_variant(self.variant, self.n))
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()


class StreamGTCRNMicro(_B200Module):
    """Drop-in for the reference `StreamGTCRNMicro` plus a native state API for many concurrent streams."""
    _stream = True

    # ---- reference-compatible call (cache tensors in the reference's own layout) -------------------
    @torch.no_grad()
    def forward(self, spec: torch.Tensor, conv_cache: torch.Tensor, tra_cache: torch.Tensor, tcn_cache: List[List[torch.Tensor]]):
        """spec [B,257,1,2]; conv_cache [2,B,16,6,33]; tra_cache [2,3,B,8,2]; tcn_cache[2][4] of [B,16,2d,33].
        Returns (spec_enh [B,257,1,2], conv_cache, tra_cache, tcn_cache) with the caches advanced by one frame."""
        self._check_spec(spec, T=1)
        B = spec.shape[0]
        assert tuple(conv_cache.shape) == (2, B, 16, 6, 33), conv_cache.shape
        assert tuple(tra_cache.shape) == (2, 3, B, 8, 2), tra_cache.shape
        for g in range(2):
            for k, d in enumerate(TCN_DILATIONS):
                assert tuple(tcn_cache[g][k].shape) == (B, 16, 2 * d, 33), tcn_cache[g][k].shape
        h = self._ensure_handle(spec.device)
        L = _lib.load()
        cc, tc = conv_cache.contiguous(), tra_cache.contiguous()
        nc = [[tcn_cache[g][k].contiguous() for k in range(4)] for g in range(2)]
        ptrs = (C.c_void_p * 8)(*[nc[g][k].data_ptr() for g in range(2) for k in range(4)])
        # The native ring-buffer state is kept between calls and re-imported only when the caller hands in cache tensors other
        # than the ones this module exported last time (or modified them through torch: `_version` moves).  The usual loop
        # (gtcrn_micro_stream.py:626-636) feeds the returned objects straight back, so it pays for one import in total.
        cached = getattr(self, "_fwd_state", None)
        key = (int(h.value), B, spec.device, tuple((t.data_ptr(), t._version) for t in [cc, tc] + [nc[g][k] for g in range(2) for k in range(4)]))
        frame = spec[:, :, 0, :].contiguous()
        out = torch.empty_like(frame)
        with torch.cuda.device(spec.device):
            s = torch.cuda.current_stream().cuda_stream
            if cached is not None and cached[0] == key:
                st = cached[1]
            else:
                st = StreamState(B, spec.device)
                _lib.check(L.gtcrnb_stream_import_caches(h, st.ptr, B, cc.data_ptr(), tc.data_ptr(), ptrs, s), "gtcrnb_stream_import_caches")
            _lib.check(L.gtcrnb_stream_step_spec(h, st.ptr, frame.data_ptr(), out.data_ptr(), B, s), "gtcrnb_stream_step_spec")
            _lib.check(L.gtcrnb_stream_export_caches(h, st.ptr, B, cc.data_ptr(), tc.data_ptr(), ptrs, s), "gtcrnb_stream_export_caches")
        if cc.data_ptr() != conv_cache.data_ptr():
            conv_cache.copy_(cc)
        if tc.data_ptr() != tra_cache.data_ptr():
            tra_cache.copy_(tc)
        new_tcn = [[nc[g][k] for k in range(4)] for g in range(2)]
        # valid for the next call only if the caller passes back exactly what we return (raw-pointer writes do not move _version)
        nkey = (int(h.value), B, spec.device, tuple((t.data_ptr(), t._version) for t in [conv_cache, tra_cache] + [new_tcn[g][k] for g in range(2) for k in range(4)]))
        self._fwd_state = (nkey, st) if (cc.data_ptr() == conv_cache.data_ptr() and tc.data_ptr() == tra_cache.data_ptr()) else None
        return out[:, :, None, :], conv_cache, tra_cache, new_tcn

    # ---- native API: state stays on the device as ring buffers ------------------------------------
    def new_state(self, n_streams: int, device=None) -> StreamState:
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._ensure_handle(dev)
        return StreamState(n_streams, dev, self._variant)

    @torch.no_grad()
    def reset(self, state: StreamState) -> None:
        h = self._ensure_handle(state.buf.device)
        with torch.cuda.device(state.buf.device):
            _lib.check(_lib.load().gtcrnb_stream_reset(h, state.ptr, state.n, torch.cuda.current_stream().cuda_stream), "gtcrnb_stream_reset")

    @torch.no_grad()
    def step(self, frame: torch.Tensor, state: StreamState, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One 16 ms hop on spectra: frame [N,257,2] contiguous -> enhanced [N,257,2]."""
        if tuple(frame.shape) != (state.n, 257, 2) or frame.dtype != torch.float32 or not frame.is_contiguous():
            raise ValueError(f"frame must be contiguous float32 [{state.n},257,2]")
        h = self._ensure_handle(frame.device)
        if out is None:
            out = torch.empty_like(frame)
        with torch.cuda.device(frame.device):
            _lib.check(_lib.load().gtcrnb_stream_step_spec(h, state.ptr, frame.data_ptr(), out.data_ptr(), state.n,
                                                           torch.cuda.current_stream().cuda_stream), "gtcrnb_stream_step_spec")
        return out

    @torch.no_grad()
    def step_wav(self, hop: torch.Tensor, state: StreamState, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One hop on waveforms: hop [N,256] new samples -> [N,256] enhanced samples delayed by 2 hops (512 samples)."""
        if tuple(hop.shape) != (state.n, 256) or hop.dtype != torch.float32 or not hop.is_contiguous():
            raise ValueError(f"hop must be contiguous float32 [{state.n},256]")
        h = self._ensure_handle(hop.device)
        if out is None:
            out = torch.empty_like(hop)
        with torch.cuda.device(hop.device):
            _lib.check(_lib.load().gtcrnb_stream_step_wav(h, state.ptr, hop.data_ptr(), out.data_ptr(), state.n,
                                                          torch.cuda.current_stream().cuda_stream), "gtcrnb_stream_step_wav")
        return out


    @torch.no_grad()
    def flush_wav(self, state: StreamState, tail: Optional[torch.Tensor] = None) -> torch.Tensor:
        """End of stream: returns [N,512], the two enhanced hops `step_wav`'s 32 ms latency still holds back, computed with the
        end-of-signal reflect padding of torch.stft -- concatenated after the `step_wav` outputs (minus the two leading zero
        hops) this is exactly `GTCRNMicro.enhance` of the whole signal.  tail [N,r], r < 256: trailing samples that did not
        fill a hop (they shape the last frame, infer.py:60-67).  The state must be reset before it is used again."""
        r = 0 if tail is None else int(tail.shape[1])
        if tail is not None and (tuple(tail.shape) != (state.n, r) or r > 255 or tail.dtype != torch.float32 or not tail.is_contiguous()):
            raise ValueError(f"tail must be contiguous float32 [{state.n}, r < 256]")
        dev = state.buf.device
        h = self._ensure_handle(dev)
        out = torch.empty((2, state.n, 256), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().gtcrnb_stream_flush_wav(h, state.ptr, tail.data_ptr() if r else None, r, out.data_ptr(), state.n,
                                                           torch.cuda.current_stream().cuda_stream), "gtcrnb_stream_flush_wav")
        return out.permute(1, 0, 2).reshape(state.n, 512)


class StreamPool:
    """N concurrent real-time streams served from HOST buffers (the frame-by-frame loop of gtcrn_micro_stream.py:618-636 for many
    callers at once).  The streams are split into groups with a device state each; per hop the H2D copy of group g + 1, the kernels
    of group g and the D2H copy of group g - 1 overlap on three CUDA streams, so only the first copy in and the last copy out are
    exposed (one serial H2D -> step -> D2H chain spends 45 % of a 160,000-stream hop in the copies).  Streams are independent, so
    the result is bit-identical to one `step_wav` over all of them."""

    GROUP_STREAMS = 8192

    def __init__(self, smodel: "StreamGTCRNMicro", n_streams: int, device=None, groups: Optional[int] = None):
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.smodel, self.n, self.device = smodel, int(n_streams), dev
        g = groups if groups is not None else max(1, min(8, self.n // self.GROUP_STREAMS))
        g = max(1, min(g, self.n))
        self.sizes = [self.n // g + (1 if i < self.n % g else 0) for i in range(g)]
        self.offsets = [sum(self.sizes[:i]) for i in range(g)]
        self.states = [smodel.new_state(k, dev) for k in self.sizes]
        self.din = [torch.empty(k, 256, device=dev) for k in self.sizes]
        self.dout = [torch.empty(k, 256, device=dev) for k in self.sizes]
        with torch.cuda.device(dev):
            self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self.ev_in = [torch.cuda.Event() for _ in range(g)]
            self.ev_step = [torch.cuda.Event() for _ in range(g)]
            self.ev_out = [torch.cuda.Event() for _ in range(g)]
        self._hops = 0

    @staticmethod
    def _check_host(t: torch.Tensor, n: int, name: str) -> None:
        if t.device.type != "cpu" or t.dtype != torch.float32 or tuple(t.shape) != (n, 256) or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous float32 CPU tensor [{n},256] (pinned for overlap)")

    @torch.no_grad()
    def step_wav_host(self, hop_in: torch.Tensor, hop_out: torch.Tensor) -> torch.Tensor:
        """One 16 ms hop for all streams: hop_in [N,256] new samples (host) -> hop_out [N,256] enhanced samples (host), delayed by
        two hops like `step_wav`.  Asynchronous: the caller's current CUDA stream completes when hop_out has landed."""
        self._check_host(hop_in, self.n, "hop_in")
        self._check_host(hop_out, self.n, "hop_out")
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            self.s_in.wait_stream(cur)                         # whatever the caller queued before this hop (e.g. a reset)
            for g, (o, k) in enumerate(zip(self.offsets, self.sizes)):
                with torch.cuda.stream(self.s_in):
                    if self._hops:
                        self.s_in.wait_event(self.ev_step[g])  # the previous hop's kernels have consumed din[g]
                    self.din[g].copy_(hop_in[o:o + k], non_blocking=True)
                    self.ev_in[g].record()
                cur.wait_event(self.ev_in[g])
                if self._hops:
                    cur.wait_event(self.ev_out[g])             # dout[g] of the previous hop is on the host
                self.smodel.step_wav(self.din[g], self.states[g], self.dout[g])
                self.ev_step[g].record(cur)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(self.ev_step[g])
                    hop_out[o:o + k].copy_(self.dout[g], non_blocking=True)
                    self.ev_out[g].record()
            cur.wait_event(self.ev_out[-1])
            for e in self.ev_out[:-1]:
                cur.wait_event(e)
        self._hops += 1
        return hop_out

    def reset(self) -> None:
        for st in self.states:
            self.smodel.reset(st)


def convert_to_stream(stream_model: nn.Module, model: nn.Module) -> None:
    """Copy offline weights into the streaming twin (same contract as convert.py:7-56, including the
    ConvTranspose2d weight re-layout and `ValueError("Key error!")` for unknown keys)."""
    src = model.state_dict()
    dst = stream_model.state_dict()
    for key in list(dst.keys()):
        if key in src:
            dst[key] = src[key]
            continue
        for infix in ("Conv1d.", "Conv2d.", ".deconv"):
            k2 = key.replace(infix, "")
            if k2 != key and k2 in src:
                dst[key] = src[k2]
                break
        else:
            k2 = key.replace("ConvTranspose2d.", "")
            if k2 == key or k2 not in src:
                raise ValueError("Key error!")
            if key.endswith("weight"):
                w = torch.flip(src[k2].permute(1, 0, 2, 3).contiguous(), dims=[-2, -1])
                assert w.shape == dst[key].shape, (w.shape, dst[key].shape)
                dst[key] = w
            else:
                dst[key] = src[k2]
    stream_model.load_state_dict(dst)


class _HostBlock:
    """Owns one gtcrnb_host_alloc block; freed when the last tensor viewing it goes away."""
    def __init__(self, ptr: int):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().gtcrnb_host_free(C.c_void_p(self.ptr))
                self.ptr = 0
        except Exception:
            pass


def pinned_empty(shape, write_combined: bool = False) -> torch.Tensor:
    """float32 (only) page-locked host tensor from the library's allocator (`gtcrnb_host_alloc[_wc]`) for `enhance_host`.
    write_combined=True is for INPUT buffers that the CPU fills sequentially and never reads: the copy engine reads them
    without cache snooping.  CPU reads of such a tensor are very slow."""
    n = 1
    for d in shape:
        n *= int(d)
    L = _lib.load()
    ptr = (L.gtcrnb_host_alloc_wc if write_combined else L.gtcrnb_host_alloc)(max(4, 4 * n))
    if not ptr:
        raise _lib.GtcrnB200Error(f"page-locked allocation of {4 * n} bytes failed")
    block = _HostBlock(int(ptr))
    buf = (C.c_float * n).from_address(block.ptr)
    buf._gtcrnb_owner = block                                   # ctypes array keeps the block alive; torch keeps the array alive
    t = torch.frombuffer(buf, dtype=torch.float32, count=n) if n else torch.empty(0)
    return t.view(*shape)


def load_state_dict_npz(path: str) -> dict:
    """state_dict saved as .npz (keys = reference state_dict keys; the fixtures under tests/golden) -> dict of tensors."""
    z = np.load(path)
    return {k: torch.from_numpy(np.array(z[k])) for k in z.files}


def load_checkpoint(path: str, stream: bool = False, map_location="cpu", unsafe: bool = False) -> nn.Module:
    """Build a model from a reference training checkpoint `{"epoch","optimizer","scheduler","model"}`
    (train.py:206-211); also tolerates a bare state_dict, `state_dict`/`model_state_dict` keys and a DDP
    `module.` prefix like the reference's export scripts (stream_onnx.py:38-47).
    The file is read with `weights_only=True` (tensors and plain containers only -- the shipped best_model_dns3.tar loads
    that way); `unsafe=True` opts into full unpickling for checkpoints that carry custom objects, and must only be used
    on trusted files."""
    ckpt = torch.load(path, map_location=map_location, weights_only=not unsafe)
    sd = ckpt
    for k in ("model", "state_dict", "model_state_dict"):
        if isinstance(ckpt, dict) and k in ckpt and isinstance(ckpt[k], dict):
            sd = ckpt[k]
            break
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
    model = GTCRNMicro()
    model.load_state_dict(sd)
    if not stream:
        return model
    smodel = StreamGTCRNMicro()
    convert_to_stream(smodel, model)
    return smodel
