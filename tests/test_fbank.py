"""Audio front end, CPU side: the numpy restatement of the Kaldi filterbank (oracle/fbank_oracle.py) against the golden
vectors minted from torchaudio (oracle/make_golden_fbank.py), and the host-side constants of vita_b200.audio_frontend."""
import os

import numpy as np
import pytest

from oracle import fbank_oracle as F

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "fbank_golden.npz"))
CASES = ["speech_1s", "noise_half_s", "silence", "short"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_torchaudio_golden(name):
    wave, ref = GOLD[name + "_wave"], GOLD[name + "_fbank"]
    got = F.fbank(wave * np.float32(32768.0))
    assert got.shape == ref.shape == (F.num_frames(wave.shape[0]), 80)
    assert np.abs(got - ref).max() <= 2e-4, "float32 log-mel features: |oracle - torchaudio| <= 2e-4"


def test_silence_hits_the_log_floor():
    got = F.fbank(np.zeros(1200, dtype=np.float32))
    assert np.all(got == np.log(np.float32(np.finfo(np.float32).eps)))


def test_frame_and_token_counts():
    import torch
    assert F.num_frames(160000) == 998 and F.num_frames(400) == 1 and F.num_frames(399) == 0
    for t in (1, 2, 3, 7, 48, 98, 400, 998, 999, 1001):
        assert F.n_llm_tokens(t) == torch.ones(t)[2::2][2::2][0::2].shape[0]     # init_model.py:57-58
    assert F.n_llm_tokens(998) == 124


def test_frontend_constants_equal_the_oracle():
    from vita_b200 import audio_frontend as A
    assert np.array_equal(A._povey_window(400), F.povey_window(400))
    banks = A._mel_banks(80, 512, 16000.0, 20.0, 0.0)
    assert np.array_equal(banks, F.mel_banks())
    nz = banks > 0
    assert nz.any(axis=1).all(), "every filter has support"
    # supports are contiguous, so a [first, last) span per filter loses nothing
    for m in range(80):
        idx = np.flatnonzero(nz[m])
        assert idx[-1] - idx[0] + 1 == idx.size
    assert A.n_llm_tokens(998) == 124
