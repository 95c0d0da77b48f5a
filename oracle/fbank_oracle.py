"""TEST INFRASTRUCTURE ONLY (see oracle/README or DESIGN.md section 1): CPU restatement of the Kaldi log-mel filterbank
the reference's audio front-end computes before the Whale encoder.

Reference call site: vita/model/multimodal_encoder/whale/init_model.py:35-60 (`audioEncoderProcessor.process`):
    waveform = waveform * (1 << 15)
    mat = kaldi.fbank(waveform, num_mel_bins=80, frame_length=25, frame_shift=10, dither=<conf>, energy_floor=0.0,
                      sample_frequency=16000)
The arithmetic lives in the third-party dependency `torchaudio.compliance.kaldi` (torchaudio is unpinned in the
reference's requirements; 2.11.0 is installed here).  Its published algorithm, restated below in float32 numpy:
  frames (snip_edges) -> subtract the frame mean -> pre-emphasis 0.97 with the first sample replicated ->
  povey window (hann ** 0.85) -> zero-pad 400 -> 512 -> |rfft|^2 -> 80 triangular mel filters (20 Hz .. Nyquist,
  mel = 1127 ln(1 + f / 700)) -> log(max(., FLT_EPSILON)).
dither must be 0 (random noise otherwise; SURVEY.md section 8 parity note 4).

Pinned against torchaudio itself: tests/golden/fbank_golden.npz, minted by oracle/make_golden_fbank.py
(tests/test_fbank.py checks |oracle - torchaudio| there).
"""
import math

import numpy as np

SAMPLE_RATE = 16000
FRAME_LEN = 400      # 25 ms
FRAME_SHIFT = 160    # 10 ms
N_FFT = 512
N_MEL = 80
PREEMPH = np.float32(0.97)
EPS = np.float32(np.finfo(np.float32).eps)   # torchaudio's EPSILON


def povey_window(n: int = FRAME_LEN) -> np.ndarray:
    """torch.hann_window(n, periodic=False) ** 0.85 (kaldi.py `_feature_window_function`, POVEY)."""
    k = np.arange(n, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * k / (n - 1))
    return (hann.astype(np.float32) ** np.float32(0.85)).astype(np.float32)


def mel_banks(n_mel: int = N_MEL, n_fft: int = N_FFT, sample_rate: float = SAMPLE_RATE, low: float = 20.0,
              high: float = 0.0) -> np.ndarray:
    """kaldi.py `get_mel_banks` without VTLN warping: [n_mel, n_fft / 2] float32."""
    f32 = np.float32
    nyquist = 0.5 * sample_rate
    if high <= 0.0:
        high += nyquist
    bin_width = sample_rate / n_fft
    mel = lambda f: f32(1127.0) * np.log(f32(1.0) + np.asarray(f, dtype=f32) / f32(700.0))
    mel_low = 1127.0 * math.log(1.0 + low / 700.0)    # python floats in the reference (mel_scale_scalar), so the
    mel_high = 1127.0 * math.log(1.0 + high / 700.0)  # tensor arithmetic below stays float32
    delta = (mel_high - mel_low) / (n_mel + 1)
    b = np.arange(n_mel, dtype=np.float32)[:, None]
    left = (mel_low + b * delta).astype(f32)
    center = (mel_low + (b + 1.0) * delta).astype(f32)
    right = (mel_low + (b + 2.0) * delta).astype(f32)
    m = mel(f32(bin_width) * np.arange(n_fft // 2, dtype=f32))[None, :]
    up = (m - left) / (center - left)
    down = (right - m) / (right - center)
    return np.maximum(f32(0.0), np.minimum(up, down)).astype(f32)


def num_frames(n_samples: int) -> int:
    return 0 if n_samples < FRAME_LEN else 1 + (n_samples - FRAME_LEN) // FRAME_SHIFT


def fbank(wave_scaled: np.ndarray) -> np.ndarray:
    """wave_scaled: [n_samples] float32, already multiplied by 2**15.  Returns [n_frames, 80] float32."""
    x = np.asarray(wave_scaled, dtype=np.float32)
    t = num_frames(x.shape[0])
    idx = np.arange(t)[:, None] * FRAME_SHIFT + np.arange(FRAME_LEN)[None, :]
    fr = x[idx]                                                       # [T, 400]
    fr = fr - fr.mean(axis=1, keepdims=True, dtype=np.float32)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)            # replicate-pad on the left
    fr = fr - PREEMPH * prev
    fr = fr * povey_window()[None, :]
    pad = np.zeros((t, N_FFT - FRAME_LEN), dtype=np.float32)
    spec = np.fft.rfft(np.concatenate([fr, pad], axis=1).astype(np.float32), axis=1)
    power = (np.abs(spec).astype(np.float32)) ** np.float32(2.0)      # [T, 257]
    banks = np.concatenate([mel_banks(), np.zeros((N_MEL, 1), dtype=np.float32)], axis=1)   # Nyquist column = 0
    e = power.astype(np.float32) @ banks.T
    return np.log(np.maximum(e.astype(np.float32), EPS)).astype(np.float32)


def n_llm_tokens(n_frames: int) -> int:
    """Length of ones(T)[2::2][2::2][0::2] (init_model.py:57-58): frames after the encoder's /4 and the adapter's /2."""
    a = len(range(2, n_frames, 2))
    b = len(range(2, a, 2))
    return len(range(0, b, 2))
