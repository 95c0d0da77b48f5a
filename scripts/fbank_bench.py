"""10 s of audio -> [998, 80] log-mel: vita_fbank on the GPU (host waveform -> device features, copy included) against
torchaudio.compliance.kaldi.fbank on the host cores (what the reference's audio_processor runs)."""
import json
import time

import numpy as np
import torch

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vita_b200.audio_frontend import AudioProcessor   # noqa: E402


def main():
    g = np.random.default_rng(0)
    wave = torch.from_numpy((0.1 * g.standard_normal(160000)).astype(np.float32)).pin_memory()
    ap = AudioProcessor("cuda")
    for _ in range(5):
        ap.process_waveform(wave, 16000)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    a.record()
    for _ in range(n):
        mat, _ = ap.process_waveform(wave, 16000)
    b.record(); b.synchronize()
    e2e_us = a.elapsed_time(b) / n * 1e3
    dev_wave = wave.cuda() * 32768.0
    from vita_b200 import ops
    a.record()
    for _ in range(n):
        ops.fbank(dev_wave, ap.window, ap.mel_t, ap.mel_span, ap.frame_len, ap.frame_shift, ap.preemph)
    b.record(); b.synchronize()
    kern_us = a.elapsed_time(b) / n * 1e3
    import torchaudio.compliance.kaldi as kaldi
    w = wave[None, :] * (1 << 15)
    kaldi.fbank(w, num_mel_bins=80, dither=0.0, energy_floor=0.0, sample_frequency=16000)
    t0 = time.perf_counter()
    for _ in range(10):
        ref = kaldi.fbank(w, num_mel_bins=80, dither=0.0, energy_floor=0.0, sample_frequency=16000)
    cpu_us = (time.perf_counter() - t0) / 10 * 1e6
    err = float((mat.cpu() - ref).abs().max())
    print(json.dumps({"frames": int(mat.shape[0]), "gpu_kernel_us": round(kern_us, 2), "gpu_host_to_features_us": round(e2e_us, 2),
                      "torchaudio_cpu_us": round(cpu_us, 1), "cpu_threads": torch.get_num_threads(), "max_abs_diff": err}))


if __name__ == "__main__":
    main()
