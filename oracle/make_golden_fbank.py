"""Mints tests/golden/fbank_golden.npz from torchaudio itself (the third-party dependency the reference calls at
whale/init_model.py:48-56).  Run in the build container: python oracle/make_golden_fbank.py"""
import os

import numpy as np
import torch
import torchaudio.compliance.kaldi as kaldi

HERE = os.path.dirname(os.path.abspath(__file__))


def synth_wave(seconds: float, seed: int) -> np.ndarray:
    """Speech-like test signal in [-1, 1]: a few gliding harmonics under a syllable-rate envelope plus a noise floor."""
    g = np.random.default_rng(seed)
    n = int(16000 * seconds)
    t = np.arange(n) / 16000.0
    f0 = 120.0 + 40.0 * np.sin(2 * np.pi * 0.7 * t)
    phase = 2 * np.pi * np.cumsum(f0) / 16000.0
    x = sum(a * np.sin(k * phase) for k, a in ((1, 0.5), (2, 0.3), (3, 0.2), (5, 0.1), (9, 0.05), (17, 0.02)))
    env = 0.55 + 0.45 * np.sin(2 * np.pi * 3.1 * t)
    x = x * env + 0.01 * g.standard_normal(n)
    return (0.5 * x / np.abs(x).max()).astype(np.float32)


def main():
    waves = {"speech_1s": synth_wave(1.0, 0), "noise_half_s": (0.1 * np.random.default_rng(1).standard_normal(8000)).astype(np.float32),
             "silence": np.zeros(1200, dtype=np.float32), "short": synth_wave(0.0252, 2)[:403]}
    out = {}
    for name, w in waves.items():
        wav = torch.from_numpy(w)[None, :] * (1 << 15)
        mat = kaldi.fbank(wav, num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0, energy_floor=0.0,
                          sample_frequency=16000)
        out[name + "_wave"] = w
        out[name + "_fbank"] = mat.numpy()
        print(name, w.shape, "->", tuple(mat.shape))
    path = os.path.join(HERE, "..", "tests", "golden", "fbank_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.normpath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
