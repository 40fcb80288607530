"""File pipeline (gtcrn_micro_b200/infer.py, the replacement for the reference's infer.py loop): host logic on CPU with a stand-in
model, and on the GPU end to end against the reference's shipped noisy -> enh example pairs."""
import os

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from gtcrn_micro_b200 import infer as I


class HalfGain:
    """Stand-in with the enhance() contract: [B,L] -> [B,256*(L//256)]."""
    def enhance(self, x):
        return 0.5 * x[:, :256 * (x.shape[1] // 256)]

    def pair_metrics(self, r, e):
        return torch.zeros(r.shape[0], 2)


def write_pair(d_noisy, d_clean, fid, n, seed, clean_n=None):
    g = np.random.default_rng(seed)
    x = (g.standard_normal(n) * 3000).astype(np.int16)
    wavfile.write(os.path.join(d_noisy, f"noisy_snr5_fileid_{fid}.wav"), 16000, x)
    wavfile.write(os.path.join(d_clean, f"clean_fileid_{fid}.wav"), 16000, x[:clean_n or n])
    return x


def test_fileid_and_job_listing(tmp_path):
    assert I.extract_fileid("/a/b/noisy_x_fileid_17.wav") == "17"
    assert I.extract_fileid("/a/b/other.wav") is None
    dn, dc = tmp_path / "n", tmp_path / "c"; dn.mkdir(); dc.mkdir()
    write_pair(dn, dc, 2, 1000, 0); write_pair(dn, dc, 1, 1000, 1)
    (dn / "notes.txt").write_text("x")
    jobs = I.list_jobs(str(dn), str(dc))
    assert [j.uid for j in jobs] == ["noisy_snr5_fileid_1", "noisy_snr5_fileid_2"]
    assert jobs[0].ref_path.endswith("clean_fileid_1.wav")
    os.remove(dc / "clean_fileid_2.wav")
    with pytest.raises(FileNotFoundError):
        I.list_jobs(str(dn), str(dc))
    wavfile.write(dc / "clean_fileid_2.wav", 16000, np.zeros(10, np.int16))
    wavfile.write(dn / "plain.wav", 16000, np.zeros(10, np.int16))
    with pytest.raises(RuntimeError):
        I.list_jobs(str(dn), str(dc))
    assert len(I.list_jobs(str(dn), None)) == 3


def test_wav_header_length_and_read(tmp_path):
    for n, dt in ((1234, np.int16), (777, np.float32)):
        p = str(tmp_path / f"a{n}.wav")
        wavfile.write(p, 16000, np.zeros(n, dt))
        assert I.wav_num_samples(p) == n
        assert len(I.read_wav(p)) == n
    p = str(tmp_path / "bad.wav")
    wavfile.write(p, 8000, np.zeros(10, np.int16))
    with pytest.raises(ValueError):
        I.read_wav(p)


def test_bucketing_and_length_fix():
    b = I.bucket_by_length([5, 7, 5, 5, 7, 9], 2)
    assert b == [[0, 2], [3], [1, 4], [5]]
    assert sorted(i for g in b for i in g) == list(range(6))
    x = torch.arange(12.0).reshape(2, 6)
    assert I.fix_length(x, 4).shape == (2, 4) and torch.equal(I.fix_length(x, 8)[:, 6:], torch.zeros(2, 2))
    q = I.to_pcm16(torch.tensor([0.0, -1e-6, 0.99999, 1.5, -2.0, 3.2 / 32768]))
    assert q.tolist() == [0, -1, 32767, 32767, -32768, 3]


def test_run_with_stand_in_model(tmp_path):
    dn, dc, de = tmp_path / "n", tmp_path / "c", tmp_path / "e"; dn.mkdir(); dc.mkdir()
    xs = {1: write_pair(dn, dc, 1, 2000, 1), 2: write_pair(dn, dc, 2, 3000, 2), 3: write_pair(dn, dc, 3, 2000, 3, clean_n=1500)}
    jobs = I.list_jobs(str(dn), str(dc))
    inf_rows, ref_rows, _ = I.run(HalfGain(), jobs, str(de), max_batch=8, device="cpu")
    assert [u for u, _ in inf_rows] == [j.uid for j in jobs] and len(ref_rows) == 3
    for fid, n_out in ((1, 2000), (2, 3000), (3, 1500)):
        fs, y = wavfile.read(de / f"noisy_snr5_fileid_{fid}_enh.wav")
        assert fs == 16000 and y.dtype == np.int16 and len(y) == n_out          # padded / truncated to the clean length
        n_valid = min(n_out, 256 * (len(xs[fid]) // 256))
        assert np.array_equal(y[:n_valid], np.floor(xs[fid][:n_valid] * 0.5).astype(np.int16))
        assert not y[n_valid:].any()
    lines = (de / "inf.scp").read_text().splitlines()
    assert lines[0].split()[0] == "noisy_snr5_fileid_1" and lines[0].split()[1].endswith("_enh.wav")
    assert len((de / "ref.scp").read_text().splitlines()) == 3


@pytest.mark.gpu
def test_pipeline_reproduces_shipped_enhanced_files(tmp_path, weights_dns3, pairs, b200_lib):
    import gtcrn_micro_b200 as G
    m = G.GTCRNMicro(); m.load_state_dict(weights_dns3)
    dn, dc, de = tmp_path / "n", tmp_path / "c", tmp_path / "e"; dn.mkdir(); dc.mkdir()
    for i in range(1, 6):
        wavfile.write(dn / f"noisy_fileid_{i}.wav", 16000, pairs[f"noisy{i}"])
        wavfile.write(dc / f"clean_fileid_{i}.wav", 16000, pairs[f"noisy{i}"])   # length donor + metric reference
    jobs = I.list_jobs(str(dn), str(dc))
    _, _, scores = I.run(m, jobs, str(de), max_batch=4, metrics=True)
    for i in range(1, 6):
        _, y = wavfile.read(de / f"noisy_fileid_{i}_enh.wav")
        enh = pairs[f"enh{i}"]
        assert len(y) == len(pairs[f"noisy{i}"])
        assert np.abs(y[:len(enh)].astype(np.int32) - enh).max() <= 1
        assert np.mean(y[:len(enh)] == enh) > 0.995
    assert len(scores) == 5 and all(np.isfinite(v).all() for v in scores.values())


@pytest.mark.gpu
def test_pair_metrics_vs_oracle(oracle, b200_lib):
    import gtcrn_micro_b200 as G
    m = G.GTCRNMicro()
    for L in (1, 255, 16000, 160000):
        x, clean = oracle.synth_batch(5, L, seed=L, return_clean=True)
        x = x + 0.01                                                             # DC offset: the metrics must centre
        out = m.pair_metrics(clean.to("cuda:0"), x.to("cuda:0")).cpu().numpy()
        for i in range(5):
            a, b = oracle.si_snr(clean[i].numpy(), x[i].numpy()), oracle.sdr(clean[i].numpy(), x[i].numpy())
            assert abs(out[i, 0] - a) <= 1e-3 * max(1.0, abs(a)), (L, i, out[i, 0], a)
            assert abs(out[i, 1] - b) <= 1e-3 * max(1.0, abs(b)), (L, i, out[i, 1], b)
    big = torch.zeros(8, 40000, device="cuda:0")
    assert m.pair_metrics(big[:, ::2], big[:, ::2]).shape == (8, 2)              # strided views
    assert m.pair_metrics(big[:0], big[:0]).shape == (0, 2)


@pytest.mark.gpu
def test_pair_metrics_vs_reference_values(metrics_golden, b200_lib):
    import gtcrn_micro_b200 as G
    m = G.GTCRNMicro()
    for i in range(4):
        r = torch.from_numpy(metrics_golden[f"ref{i}"])[None].to("cuda:0")
        x = torch.from_numpy(metrics_golden[f"inf{i}"])[None].to("cuda:0")
        out = m.pair_metrics(r, x).cpu().numpy()[0]
        assert abs(out[0] - float(metrics_golden[f"si_snr{i}"])) <= 1e-3, (i, out)
        assert abs(out[1] - float(metrics_golden[f"sdr{i}"])) <= 1e-3, (i, out)


@pytest.mark.gpu
def test_pcm16_conversions_and_float_wav_files(tmp_path, weights_dns3, pairs, b200_lib):
    import gtcrn_micro_b200 as G
    m = G.GTCRNMicro(); m.load_state_dict(weights_dns3)
    g = torch.Generator().manual_seed(1)
    for n in (1, 7, 8, 1000, 40003):
        q = torch.randint(-32768, 32768, (n,), generator=g, dtype=torch.int32).to(torch.int16).to("cuda:0")
        f = m.pcm16_to_float(q)
        assert torch.equal(f.cpu(), q.cpu().float() / 32768.0)
        assert torch.equal(m.float_to_pcm16(f), q)                                   # exact round trip
        x = (torch.randn(n, generator=g) * 0.7).to("cuda:0")
        want = torch.floor(x * 32768.0).clamp_(-32768, 32767).to(torch.int16)
        assert torch.equal(m.float_to_pcm16(x), want)
        assert torch.equal(m.float_to_pcm16(x[1:]), want[1:])                        # unaligned view -> scalar tail path
    # float32 wav files take the float branch of the pipeline and must give the same PCM16 output as PCM16 files
    dn, de1, de2 = tmp_path / "n16", tmp_path / "e16", tmp_path / "e32"; dn.mkdir()
    dn32 = tmp_path / "n32"; dn32.mkdir()
    for i in (1, 2):
        wavfile.write(dn / f"noisy_fileid_{i}.wav", 16000, pairs[f"noisy{i}"])
        wavfile.write(dn32 / f"noisy_fileid_{i}.wav", 16000, pairs[f"noisy{i}"].astype(np.float32) / 32768.0)
    I.run(m, I.list_jobs(str(dn), None), str(de1), max_batch=4)
    I.run(m, I.list_jobs(str(dn32), None), str(de2), max_batch=4)
    for i in (1, 2):
        a = wavfile.read(de1 / f"noisy_fileid_{i}_enh.wav")[1]; b = wavfile.read(de2 / f"noisy_fileid_{i}_enh.wav")[1]
        assert a.dtype == np.int16 and np.array_equal(a, b)
