"""Audio front end on the GPU: waveform -> kaldi log-mel features for the Whale encoder.

Host-side mirror of the reference's `audioEncoderProcessor` (vita/model/multimodal_encoder/whale/init_model.py:28-60):
same `process(wav_path) -> (fbank [T, 80], n_llm_tokens)` contract, but the 25 ms / 10 ms / 80-bin filterbank runs in
`vita_fbank` (csrc/fbank.cu) so the features are produced where the encoder consumes them.  File decoding and sample-
rate conversion stay the reference's own torchaudio CPU code (out of scope).  `dither`
must be 0 for a reproducible path (the reference's training config uses 1.0; inference parity needs 0 -- SURVEY.md
section 8 parity note 4).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops

SAMPLE_RATE = 16000


def _povey_window(n: int) -> np.ndarray:
    k = np.arange(n, dtype=np.float64)
    hann = (0.5 - 0.5 * np.cos(2.0 * np.pi * k / (n - 1))).astype(np.float32)
    return (hann ** np.float32(0.85)).astype(np.float32)


def _mel_banks(n_mel: int, n_fft: int, sample_rate: float, low: float, high: float) -> np.ndarray:
    """Triangular filters on the mel scale 1127 ln(1 + f / 700), Kaldi style: [n_mel, n_fft / 2] float32."""
    f32 = np.float32
    nyquist = 0.5 * sample_rate
    if high <= 0.0:
        high += nyquist
    mel_low = 1127.0 * math.log(1.0 + low / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high / 700.0)
    delta = (mel_high - mel_low) / (n_mel + 1)
    b = np.arange(n_mel, dtype=f32)[:, None]
    left, center, right = [(mel_low + (b + o) * delta).astype(f32) for o in (0.0, 1.0, 2.0)]
    freqs = f32(sample_rate / n_fft) * np.arange(n_fft // 2, dtype=f32)
    m = (f32(1127.0) * np.log(f32(1.0) + freqs / f32(700.0)))[None, :]
    return np.maximum(f32(0.0), np.minimum((m - left) / (center - left), (right - m) / (right - center))).astype(f32)


def n_llm_tokens(n_frames: int) -> int:
    """Tokens the LLM sees for n_frames of fbank: len(ones(T)[2::2][2::2][0::2]) (init_model.py:57-58)."""
    a = len(range(2, n_frames, 2))
    b = len(range(2, a, 2))
    return len(range(0, b, 2))


class AudioProcessor:
    """`audio_processor` of the audio tower: `.process(path)` as in the reference, `.process_waveform(wave, sr)` for
    callers that already hold samples."""

    def __init__(self, device="cuda", num_mel_bins: int = 80, frame_length_ms: float = 25.0, frame_shift_ms: float = 10.0,
                 dither: float = 0.0, sample_rate: int = SAMPLE_RATE, strict_reference: bool = False):
        self.strict_reference = strict_reference
        if dither != 0.0:
            raise ValueError("vita_b200 computes the deterministic filterbank only (dither must be 0.0)")
        self.device = torch.device(device)
        self.sample_rate = sample_rate
        self.frame_len = int(sample_rate * frame_length_ms * 0.001)
        self.frame_shift = int(sample_rate * frame_shift_ms * 0.001)
        n_fft = 1 << (self.frame_len - 1).bit_length()
        if n_fft != 512:
            raise ValueError("vita_fbank is built for a 512-point spectrum (25 ms frames at 16 kHz)")
        banks = _mel_banks(num_mel_bins, n_fft, float(sample_rate), 20.0, 0.0)
        nz = banks > 0
        span = np.stack([nz.argmax(axis=1), banks.shape[1] - nz[:, ::-1].argmax(axis=1)], axis=1).astype(np.int32)
        self.window = torch.from_numpy(_povey_window(self.frame_len)).to(self.device)
        self.mel_t = torch.from_numpy(np.ascontiguousarray(banks.T)).to(self.device)     # [256, n_mel]
        self.mel_span = torch.from_numpy(span).to(self.device)
        self.preemph = 0.97

    def process_waveform(self, waveform: torch.Tensor, sample_rate: int = SAMPLE_RATE):
        """waveform: [channels, n] or [n] in [-1, 1] (torchaudio.load convention); channel 0 is used."""
        if waveform.dim() == 2:
            waveform = waveform[0]
        if sample_rate != self.sample_rate:
            # DEVIATION (documented in DESIGN.md): the reference resamples to 16 kHz (init_model.py:41-45) but then
            # calls kaldi.fbank with sample_frequency = the ORIGINAL rate (:48-56), so its window length and mel banks
            # follow the original rate (e.g. 1102-sample frames / 2048-point spectrum for 44.1 kHz).  vita_fbank is the
            # 25 ms / 512-point filterbank of 16 kHz audio; the features of non-16 kHz files therefore differ from the
            # reference's.  `strict_reference=True` refuses instead of deviating.
            if self.strict_reference:
                raise NotImplementedError(
                    f"sample rate {sample_rate}: the reference derives the filterbank from the original rate "
                    "(whale/init_model.py:48-56); vita_fbank implements the 16 kHz geometry only")
            import warnings
            warnings.warn(f"audio at {sample_rate} Hz is resampled to {self.sample_rate} Hz and featurised with the "
                          "16 kHz filterbank; the reference would use window / mel banks of the original rate "
                          "(whale/init_model.py:48-56)", stacklevel=2)
            import torchaudio  # reference behaviour: CPU resampler (init_model.py:41-45)
            waveform = torchaudio.transforms.Resample(orig_freq=sample_rate, new_freq=self.sample_rate)(waveform.cpu().float())
        wave = waveform.to(self.device, torch.float32).contiguous() * float(1 << 15)
        if wave.numel() < self.frame_len:
            raise AssertionError(f"choose a window size {self.frame_len} that is [2, {wave.numel()}]")   # kaldi.py's assert
        mat = ops.fbank(wave, self.window, self.mel_t, self.mel_span, self.frame_len, self.frame_shift, self.preemph)
        return mat, n_llm_tokens(mat.shape[0])

    def process(self, wav_path: str):
        import torchaudio
        waveform, sample_rate = torchaudio.load(wav_path)
        return self.process_waveform(waveform, sample_rate)
