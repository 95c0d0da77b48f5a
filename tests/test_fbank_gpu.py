"""vita_fbank (csrc/fbank.cu) through the reference-shaped `audio_processor` against the oracle and against the
torchaudio golden vectors.  Tolerance: float32 log-mel features, |delta| <= 1e-3 (radix-2 FFT vs pocketfft rounding;
measured ~1e-4)."""
import os

import numpy as np
import pytest
import torch

from oracle import fbank_oracle as F

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "fbank_golden.npz"))
TOL = 1e-3


@pytest.fixture(scope="module")
def proc():
    from vita_b200.audio_frontend import AudioProcessor
    return AudioProcessor("cuda")


@pytest.mark.parametrize("name", ["speech_1s", "noise_half_s", "silence", "short"])
def test_kernel_matches_torchaudio_golden_and_oracle(proc, name):
    wave, ref = GOLD[name + "_wave"], GOLD[name + "_fbank"]
    mat, n_tok = proc.process_waveform(torch.from_numpy(wave)[None, :], 16000)
    assert mat.is_cuda and mat.dtype == torch.float32 and tuple(mat.shape) == ref.shape
    got = mat.cpu().numpy()
    assert np.abs(got - ref).max() <= TOL, f"vs torchaudio: {np.abs(got - ref).max()}"
    assert np.abs(got - F.fbank(wave * np.float32(32768.0))).max() <= TOL
    assert n_tok == F.n_llm_tokens(ref.shape[0])


def test_ten_seconds_full_size(proc):
    """BASELINE configs[2] audio length: 10 s -> 998 frames -> 124 LLM tokens."""
    g = np.random.default_rng(5)
    t = np.arange(160000) / 16000.0
    wave = (0.3 * np.sin(2 * np.pi * (200 + 150 * np.sin(2 * np.pi * 0.5 * t)) * t) + 0.05 * g.standard_normal(160000)).astype(np.float32)
    mat, n_tok = proc.process_waveform(torch.from_numpy(wave), 16000)
    assert tuple(mat.shape) == (998, 80) and n_tok == 124
    want = F.fbank(wave * np.float32(32768.0))
    err = np.abs(mat.cpu().numpy() - want)
    assert err.max() <= TOL, f"max {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


def test_edges(proc):
    one = torch.from_numpy(GOLD["speech_1s_wave"][:400].copy())
    mat, n_tok = proc.process_waveform(one, 16000)
    assert tuple(mat.shape) == (1, 80) and n_tok == 0
    with pytest.raises(AssertionError):
        proc.process_waveform(one[:399], 16000)     # kaldi.fbank asserts window_size <= len(waveform)
    sil, _ = proc.process_waveform(torch.zeros(1200), 16000)
    assert torch.all(sil == float(np.log(np.float32(np.finfo(np.float32).eps))))


def test_features_feed_the_encoder(proc):
    """The processor's output is what encode_audios consumes (fp32 [T, 80] on the device, no host round trip)."""
    from vita_b200 import weights as W
    from vita_b200.config import VitaConfig
    from vita_b200.model.vita_mixtral import VITAMixtralForCausalLM
    cfg = VitaConfig.tiny()
    packed = W.pack(W.synthetic_state(cfg, 0), cfg, "cuda") if hasattr(W, "pack") else None
    if packed is None:
        pytest.skip("no pack helper")
    model = VITAMixtralForCausalLM(cfg, packed, "cuda", max_seq_len=256, max_new_tokens=8)
    ap = model.get_audio_encoder().audio_processor
    mat, n_tok = ap.process_waveform(torch.from_numpy(GOLD["speech_1s_wave"]), 16000)
    out = model.encode_audios(mat[None], torch.tensor([mat.shape[0]]))
    assert out["inputs_embeds"].shape[1] == n_tok


def test_other_sample_rates_warn_or_refuse(proc):
    """whale/init_model.py:41-56 featurises non-16 kHz audio with the ORIGINAL rate's window / mel banks; vita_fbank is
    the 16 kHz geometry: the deviation is announced (warning) or refused (strict_reference)."""
    from vita_b200.audio_frontend import AudioProcessor
    wave = torch.randn(44100) * 0.1
    with pytest.warns(UserWarning, match="original rate"):
        mat, n_tok = proc.process_waveform(wave, 44100)
    assert mat.shape[1] == 80 and abs(mat.shape[0] - 98) <= 1          # 1 s -> ~98 frames at 16 kHz
    strict = AudioProcessor("cuda", strict_reference=True)
    with pytest.raises(NotImplementedError, match="original rate"):
        strict.process_waveform(wave, 44100)
