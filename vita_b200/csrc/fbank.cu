// Kaldi-compatible log-mel filterbank on the GPU: the CPU stage in front of the Whale encoder
// (reference: vita/model/multimodal_encoder/whale/init_model.py:48-56 -> torchaudio.compliance.kaldi.fbank).
//
// One CTA per frame (256 threads): frame -> DC removal -> pre-emphasis -> povey window -> 512-point radix-2 FFT in
// shared memory -> power spectrum -> triangular mel filters (only their non-zero span) -> log.  10 s of audio are 998
// frames x 400 samples: the kernel is launch/latency sized (HBM traffic 0.64 MB in, 0.32 MB out), what matters is that
// the waveform goes host -> device once and the features never visit the host.
#include <cfloat>

#include "common.h"
#include "ptx.cuh"

namespace vita {

constexpr int FB_NFFT = 512;
constexpr int FB_THREADS = 256;

__device__ __forceinline__ int fb_bitrev9(int i) { return static_cast<int>(__brev(static_cast<unsigned>(i)) >> 23); }

__global__ void __launch_bounds__(FB_THREADS)
fbank_kernel(const float* __restrict__ wave, const float* __restrict__ window, const float* __restrict__ mel_t,
             const int2* __restrict__ mel_span, float* __restrict__ out, int frame_len, int frame_shift, int n_mel,
             float preemph) {
    __shared__ float s_re[FB_NFFT], s_im[FB_NFFT];
    __shared__ float s_tw_c[FB_NFFT / 2], s_tw_s[FB_NFFT / 2];
    __shared__ float s_x[FB_NFFT];
    __shared__ float s_red[FB_THREADS / 32];
    const int f = blockIdx.x, t = threadIdx.x;
    const float* x = wave + static_cast<long long>(f) * frame_shift;

    // twiddles exp(-2 pi i k / 512), k < 256
    {
        float s, c;
        sincospif(-static_cast<float>(t) / 256.0f, &s, &c);
        s_tw_c[t] = c;
        s_tw_s[t] = s;
    }
    // frame mean
    float part = 0.0f;
    for (int i = t; i < frame_len; i += FB_THREADS) {
        const float v = x[i];
        s_x[i] = v;
        part += v;
    }
    part = warp_sum(part);
    if ((t & 31) == 0) s_red[t >> 5] = part;
    __syncthreads();
    float mean = 0.0f;
#pragma unroll
    for (int w = 0; w < FB_THREADS / 32; ++w) mean += s_red[w];
    mean /= static_cast<float>(frame_len);
    // DC removal, pre-emphasis (first sample replicated), window, zero padding; bit-reversed placement for the DIT FFT
    for (int i = t; i < FB_NFFT; i += FB_THREADS) {
        float v = 0.0f;
        if (i < frame_len) {
            const float cur = s_x[i] - mean;
            const float prev = s_x[i > 0 ? i - 1 : 0] - mean;
            v = (cur - preemph * prev) * window[i];
        }
        const int r = fb_bitrev9(i);
        s_re[r] = v;
        s_im[r] = 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 9; ++s) {
        const int half = 1 << s;
        const int pos = t & (half - 1);
        const int i0 = ((t >> s) << (s + 1)) + pos, i1 = i0 + half;
        const int k = pos << (8 - s);
        const float c = s_tw_c[k], sn = s_tw_s[k];
        const float br = s_re[i1] * c - s_im[i1] * sn, bi = s_re[i1] * sn + s_im[i1] * c;
        const float ar = s_re[i0], ai = s_im[i0];
        s_re[i0] = ar + br; s_im[i0] = ai + bi;
        s_re[i1] = ar - br; s_im[i1] = ai - bi;
        __syncthreads();
    }
    // power spectrum of bins 0..255 (the Nyquist bin has zero weight in every filter)
    s_x[t] = s_re[t] * s_re[t] + s_im[t] * s_im[t];
    __syncthreads();
    if (t < n_mel) {
        const int2 span = mel_span[t];   // [first, last) non-zero fft bins of filter t
        float e = 0.0f;
        for (int k = span.x; k < span.y; ++k) e += s_x[k] * mel_t[k * n_mel + t];
        out[static_cast<long long>(f) * n_mel + t] = logf(fmaxf(e, FLT_EPSILON));
    }
}

}  // namespace vita

using namespace vita;

extern "C" int vita_fbank(const float* wave, int64_t n_samples, const float* window, const float* mel_weights_t,
                          const int32_t* mel_span, float* out, int64_t frame_len, int64_t frame_shift, int64_t n_mel,
                          float preemph, void* stream) {
    VITA_REQUIRE(frame_len >= 2 && frame_len <= FB_NFFT && frame_shift > 0, "fbank: 2 <= frame_len <= 512, frame_shift > 0");
    VITA_REQUIRE(n_mel >= 1 && n_mel <= FB_THREADS, "fbank: 1 <= n_mel <= 256");
    VITA_REQUIRE(wave && window && mel_weights_t && mel_span && out, "fbank: null pointer");
    if (n_samples < frame_len) return VITA_OK;   // no complete frame (snip_edges)
    const int64_t n_frames = 1 + (n_samples - frame_len) / frame_shift;
    fbank_kernel<<<static_cast<unsigned>(n_frames), FB_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        wave, window, mel_weights_t, reinterpret_cast<const int2*>(mel_span), out, static_cast<int>(frame_len),
        static_cast<int>(frame_shift), static_cast<int>(n_mel), preemph);
    return check_launch("fbank_kernel");
}
