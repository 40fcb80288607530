"""Pin the oracle against the real reference (dev container only: needs /root/reference).

    python oracle/check_against_reference.py

Imports the reference's own GTCRNMicro / StreamGTCRNMicro / convert_to_stream (nothing is copied) and
checks every oracle function against them on seeded inputs, with the shipped checkpoint and with a
randomised weight set, plus the five shipped noisy->enh golden pairs.  Prints one line per check and
exits non-zero on the first failure.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("GTCRN_REFERENCE", "/root/reference")

import gtcrn_oracle as O  # noqa: E402


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF} (this script only runs in the dev container)")
    sys.path.insert(0, REF)
    # gtcrn_micro_stream.py:8 imports soundfile at module top but only __main__ uses it
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    from gtcrn_micro.models.gtcrn_micro import GTCRNMicro
    from gtcrn_micro.streaming.conversion.convert import convert_to_stream
    from gtcrn_micro.streaming.gtcrn_micro_stream import StreamGTCRNMicro
    return GTCRNMicro, StreamGTCRNMicro, convert_to_stream


def reference_metric_functions():
    """The reference's own sisnr_metric / sdr_metric (eval/eval_intrusive_metrics.py:74-94), compiled from its source file in
    isolation: the module itself imports pesq / pystoi / p_tqdm, which this image does not have."""
    import ast
    import numpy as np
    path = os.path.join(REF, "gtcrn_micro/eval/eval_intrusive_metrics.py")
    tree = ast.parse(open(path).read(), filename=path)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("sisnr_metric", "sdr_metric"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns["sisnr_metric"], ns["sdr_metric"]


def randomise(model, seed):
    """Random weights with non-trivial BN statistics and PReLU slopes (ERB buffers untouched)."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    for k, v in sd.items():
        if k.startswith("erb."):
            continue
        if k.endswith("running_var"):
            sd[k] = 0.2 + 2.0 * torch.rand(v.shape, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.3 * torch.randn(v.shape, generator=g)
        elif k.endswith("num_batches_tracked"):
            continue
        elif ".bn" in k or "_bn" in k:
            sd[k] = (1.0 + 0.3 * torch.randn(v.shape, generator=g)) if k.endswith("weight") else 0.2 * torch.randn(v.shape, generator=g)
        elif "act" in k:
            sd[k] = torch.full(v.shape, 0.25) + 0.3 * torch.randn(v.shape, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
        else:
            fan_in = max(1, v[0].numel())
            sd[k] = torch.randn(v.shape, generator=g) / np.sqrt(fan_in)
    model.load_state_dict(sd)


def check(name, err, tol):
    ok = err <= tol
    print(f"{'ok  ' if ok else 'FAIL'} {name:58s} max|diff| = {err:.3e}  (tol {tol:.1e})")
    if not ok:
        raise SystemExit(1)


def read_wav_int16(path):
    from scipy.io import wavfile
    fs, x = wavfile.read(path)
    assert fs == 16000 and x.dtype == np.int16 and x.ndim == 1
    return x


def main():
    torch.manual_seed(0)
    GTCRNMicro, StreamGTCRNMicro, convert_to_stream = import_reference()
    ckpt = torch.load(os.path.join(REF, "gtcrn_micro/ckpts/best_model_dns3.tar"), map_location="cpu")
    win = torch.hann_window(512).pow(0.5)

    # --- evaluation metrics (on-device SI-SNR / SDR are checked against these restatements)
    import numpy as np
    sisnr_ref, sdr_ref = reference_metric_functions()
    g = np.random.default_rng(5)
    for n, noise in ((16000, 0.3), (160000, 0.05), (777, 1.5)):
        r = g.standard_normal(n) * 0.1
        x = 0.8 * r + noise * 0.1 * g.standard_normal(n) + 0.01
        check(f"si_snr n={n}", abs(sisnr_ref(r, x) - O.si_snr(r, x)), 1e-9)
        check(f"sdr n={n}", abs(sdr_ref(r, x) - O.sdr(r, x)), 1e-9)

    # --- ERB bank == checkpoint buffers
    bank = torch.from_numpy(O.erb_filter_bank())
    check("erb_filter_bank vs ckpt erb_fc.weight", (bank - ckpt["model"]["erb.erb_fc.weight"]).abs().max().item(), 0.0)
    check("erb_filter_bank.T vs ckpt ierb_fc.weight", (bank.T - ckpt["model"]["erb.ierb_fc.weight"]).abs().max().item(), 0.0)

    for tag, seed in (("ckpt", None), ("rand", 7)):
        model = GTCRNMicro().eval()
        if seed is None:
            model.load_state_dict(ckpt["model"])
        else:
            randomise(model, seed)
        smodel = StreamGTCRNMicro().eval()
        convert_to_stream(smodel, model)
        W = {k: v.clone() for k, v in model.state_dict().items()}
        Ws = O.stream_to_offline_keys({k: v.clone() for k, v in smodel.state_dict().items()})
        for k in W:
            assert torch.equal(W[k], Ws[k]), k

        x = 0.1 * torch.randn(2, 256 * 40 + 77)
        with torch.no_grad():
            # STFT / iSTFT
            s_ref = torch.stft(x, 512, 256, 512, win, return_complex=False)
            s_or = O.stft(x)
            check(f"[{tag}] stft", (s_ref - s_or).abs().max().item(), 2e-5)
            y_ref = model(s_ref)
            y_or = O.forward_spec(W, s_ref)
            scale = max(1.0, y_ref.abs().max().item())
            check(f"[{tag}] forward_spec (offline, rel to max {scale:.1f})", (y_ref - y_or).abs().max().item() / scale, 2e-5)
            w_ref = torch.istft(torch.view_as_complex(y_ref.contiguous()), 512, 256, 512, win)
            w_or = O.istft(y_ref)
            check(f"[{tag}] istft", (w_ref - w_or).abs().max().item(), 1e-6)
            check(f"[{tag}] enhance (wav->wav)", (w_ref - O.enhance(W, x)).abs().max().item(), 2e-6)

            # streaming: reference frame loop vs oracle, including cache layout conversion
            B = x.shape[0]
            conv_cache = torch.zeros(2, B, 16, 6, 33)
            tra_cache = torch.zeros(2, 3, B, 8, 2)
            tcn_cache = [[torch.zeros(B, 16, 2 * d, 33) for d in (1, 2, 4, 8)] for _ in range(2)]
            st = O.StreamState(B)
            ys_ref, ys_or = [], []
            for t in range(s_ref.shape[2]):
                yi, conv_cache, tra_cache, tcn_cache = smodel(s_ref[:, :, t:t + 1], conv_cache, tra_cache, tcn_cache)
                ys_ref.append(yi)
                ys_or.append(O.forward_spec(W, s_ref[:, :, t:t + 1], st))
                if t == 20:
                    cc, tc, nc = st.to_reference()
                    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp(min=1.0)).item()
                    e = max(rel(cc, conv_cache), rel(tc, tra_cache),
                            max(rel(a, b) for g in range(2) for a, b in zip(nc[g], tcn_cache[g])))
                    check(f"[{tag}] stream caches after 21 frames (reference layout, rel)", e, 2e-5)
                    st2 = O.StreamState.from_reference(cc, tc, nc)
                    assert all(torch.equal(a, b) for a, b in zip(st2.gt, st.gt))
            ys_ref = torch.cat(ys_ref, dim=2)
            ys_or = torch.cat(ys_or, dim=2)
            check(f"[{tag}] stream frame loop vs reference stream", (ys_ref - ys_or).abs().max().item() / scale, 2e-5)
            check(f"[{tag}] oracle stream vs oracle offline", (ys_or - y_or).abs().max().item() / scale, 2e-5)
            # chunked continuation (T = 7 per call) through the same state object
            st = O.StreamState(B)
            yc = torch.cat([O.forward_spec(W, s_ref[:, :, t:t + 7], st) for t in range(0, s_ref.shape[2], 7)], dim=2)
            check(f"[{tag}] oracle chunked(7) vs offline", (yc - y_or).abs().max().item() / scale, 2e-5)

    # --- golden pairs shipped with the reference (full 31 s files)
    model = GTCRNMicro().eval()
    model.load_state_dict(ckpt["model"])
    W = {k: v.clone() for k, v in model.state_dict().items()}
    for i in range(1, 6):
        noisy = read_wav_int16(os.path.join(REF, f"gtcrn_micro/examples/gtcrn_micro/noisy{i}.wav"))
        enh = read_wav_int16(os.path.join(REF, f"gtcrn_micro/examples/gtcrn_micro/enh{i}.wav"))
        x = torch.from_numpy(noisy.astype(np.float32) / 32768.0)[None]
        out = O.enhance(W, x)[0].numpy()
        n = min(len(out), len(enh))
        err = np.abs(out[:n] - enh[:n].astype(np.float32) / 32768.0).max()
        exact = np.mean(np.floor(out[:n] * 32768.0) == enh[:n])
        check(f"golden pair {i}: oracle vs enh{i}.wav ({100 * exact:.3f}% floor-exact)", float(err), 3.1e-5)
    print("all oracle checks passed")


if __name__ == "__main__":
    main()
