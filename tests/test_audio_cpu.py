"""CPU: AudioSignal subset (resample / loudness / normalise / peak clamp / wav io).  audiotools is absent, so
these are self-consistency and known-answer checks (BS.1770: a 997 Hz full-scale sine reads -3.01 LUFS)."""
import math
import os

import numpy as np
import torch

from vampnet_b200.audio import AudioSignal


def _sine(freq, sr, secs, amp=1.0):
    t = torch.arange(int(sr * secs)) / sr
    return amp * torch.sin(2 * math.pi * freq * t)


def test_loudness_known_answer():
    s = AudioSignal(_sine(997.0, 48000, 5.0), 48000)
    assert abs(float(s.loudness()) - (-3.01)) < 0.1
    s2 = AudioSignal(_sine(997.0, 44100, 5.0, amp=0.1), 44100)
    assert abs(float(s2.loudness()) - (-23.01)) < 0.1


def test_normalize_and_peak():
    s = AudioSignal(_sine(440.0, 44100, 3.0, amp=0.05), 44100).normalize(-24.0)
    assert abs(float(s.loudness()) + 24.0) < 0.05
    loud = AudioSignal(_sine(440.0, 44100, 1.0, amp=3.0), 44100).ensure_max_of_audio(1.0)
    assert abs(float(loud.samples.abs().max()) - 1.0) < 1e-5


def test_resample_preserves_tone_and_length():
    s = AudioSignal(_sine(1000.0, 48000, 1.0), 48000).resample(44100)
    assert s.sample_rate == 44100 and s.signal_length == 44100
    ref = _sine(1000.0, 44100, 1.0)
    assert (s.samples[0, 0, 2000:-2000] - ref[2000:-2000]).abs().max() < 2e-3
    assert AudioSignal(ref, 44100).resample(44100).signal_length == 44100


def test_to_mono_clone_and_wav_roundtrip(tmp_path):
    x = torch.stack([_sine(300.0, 22050, 0.5, 0.5), _sine(500.0, 22050, 0.5, 0.25)])
    s = AudioSignal(x, 22050)
    assert s.batch_size == 1 and s.num_channels == 2
    c = s.clone().to_mono()
    assert c.num_channels == 1 and s.num_channels == 2
    p = os.path.join(tmp_path, "a.wav")
    s.write(p)
    r = AudioSignal(p)
    assert r.sample_rate == 22050 and r.samples.shape == s.samples.shape
    assert (r.samples - s.samples).abs().max() < 1e-4 + 1 / 32768


def test_loudness_against_bs1770_table_coefficients():
    """Batched, frequency-domain, any-sample-rate loudness of the product vs the time-domain restatement with the
    coefficients printed in BS.1770-4 for 48 kHz (oracle/audio_oracle.py): programme material with pauses, so both
    gates act; a batch with a silent item and a quiet item; stereo."""
    from oracle import audio_oracle as ao
    g = torch.Generator().manual_seed(0)
    sr, n = 48000, 48000 * 4
    x = torch.randn(4, 2, n, generator=g)
    env = torch.ones(n)
    env[sr:2 * sr] = 0.003            # one second close to silence: below the relative gate
    x = x * env * torch.tensor([0.2, 0.02, 1e-7, 0.5])[:, None, None]
    x[3, 1] *= 0.1                    # unbalanced stereo
    x[0] = torch.cumsum(x[0], -1) * 0.02   # low-frequency heavy: the RLB high-pass matters
    got = AudioSignal(x, sr).loudness()
    want = [ao.integrated_loudness_48k(x[i].numpy()) for i in range(4)]
    for i in range(4):
        assert abs(float(got[i]) - want[i]) < 0.05, (i, float(got[i]), want[i])
    assert want[2] == -70.0 and float(got[2]) == -70.0


def test_resample_against_scipy_polyphase():
    from oracle import audio_oracle as ao
    g = torch.Generator().manual_seed(1)
    sr_in, sr_out, n = 48000, 44100, 48000
    # band-limited test signal (a few tones below 15 kHz + low-passed noise): inside both designs' pass bands
    t = torch.arange(n) / sr_in
    x = sum(a * torch.sin(2 * math.pi * f * t + p) for f, a, p in ((220.0, 0.3, 0.1), (3300.0, 0.2, 1.0), (14000.0, 0.1, 2.0)))
    noise = torch.randn(n, generator=g)
    spec = torch.fft.rfft(noise)
    spec[int(12000 / (sr_in / 2) * (n // 2)):] = 0
    x = x + 0.05 * torch.fft.irfft(spec, n=n)
    got = AudioSignal(x[None, None], sr_in).resample(sr_out).samples[0, 0].numpy()
    want = ao.resample_poly(x.numpy(), sr_in, sr_out)
    assert got.shape == want.shape == (44100,)
    err = np.abs(got - want)[2000:-2000]
    assert err.max() < 3e-3, err.max()
