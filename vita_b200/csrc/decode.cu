// Greedy-decode step kernels (one new token per sequence): weight-streaming GEMVs, HBM-bound.
//
// Every linear of the decode step is a matrix-vector product whose cost is reading the weight once (25.2 GB per
// token for Mixtral-8x7B).  All of them run on one skeleton, stream_gemv_kernel<Op>:
//   * one CTA per SM, each owning a contiguous, balanced range of work items (an item = two weight rows);
//   * warp 8 is a producer: one thread issues cp.async.bulk (TMA 1-D) copies of the item's rows into an 8-stage,
//     16 KB/stage shared-memory ring with mbarrier complete_tx, so ~128 KB per SM is always in flight no matter what
//     the consumers are doing;
//   * warps 0-7 are consumers: conflict-free 16-byte LDS of weights and of the activation vector (bf16 in smem),
//     fp32 FMA, warp-shuffle + one named barrier per item, then a fused per-item epilogue (RoPE + paged-KV write,
//     residual add, SiLU*up, routing-weighted combine, bf16 logits + packed atomic arg-max).
// Ops: QKV (input RMSNorm fused), O-proj (+residual), router (separate tiny kernel), expert gate/up (2 selected
// experts), expert down (+weighted combine +residual), LM head (+final RMSNorm, +argmax).
//
// Reference: the decode step of HF generate() over transformers MixtralDecoderLayer (vita_mixtral.py:158-173,
// video_audio_demo.py:257-270); vLLM twin web_demo/vllm_tools/vllm_file/mixtral.py:491-566.
#include "common.h"
#include "ptx.cuh"

namespace vita {

constexpr int GV_KC = 4096;                    // elements per row per stage
constexpr int GV_STAGE_BYTES = 2 * GV_KC * 2;  // two rows of bf16
constexpr int GV_STAGES = 8;
constexpr int GV_GROUP_WARPS = 8;              // consumer warps per group
constexpr int GV_CONSUMERS = 512;              // two groups of 8 warps working on alternating batches of stages
constexpr int GV_THREADS = GV_CONSUMERS + 32;  // + producer warp
constexpr int GV_NB = 4;                       // items per batch (one reduction + one named barrier per batch)
constexpr int GV_PREP_FLOATS = 256;            // per-CTA epilogue constants prefetched before streaming starts

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_barrier_all() { asm volatile("bar.sync 3, 512;" ::: "memory"); }
__device__ __forceinline__ void group_barrier(int group) {
    asm volatile("bar.sync %0, 256;" ::"r"(group + 1) : "memory");
}

// sum over the 512 consumer threads; scratch: >= 16 floats
__device__ __forceinline__ float consumer_sum(float v, float* scratch) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    consumer_barrier_all();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += scratch[i];
    consumer_barrier_all();
    return t;
}

// xs[0..K) = bf16(rmsnorm(h) * w)   (consumer threads only)
__device__ __forceinline__ void load_x_rmsnorm(const __nv_bfloat16* h, const __nv_bfloat16* w, __nv_bfloat16* xs, int K,
                                               float eps, float* scratch) {
    float ss = 0.0f;
    for (int i = threadIdx.x * 8; i < K; i += GV_CONSUMERS * 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(h + i);
        const uint32_t a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) ss += bf16_lo(a[e]) * bf16_lo(a[e]) + bf16_hi(a[e]) * bf16_hi(a[e]);
    }
    const float tot = consumer_sum(ss, scratch);
    const float inv = rsqrtf(tot / static_cast<float>(K) + eps);
    for (int i = threadIdx.x * 8; i < K; i += GV_CONSUMERS * 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(h + i);
        const uint4 g = __ldg(reinterpret_cast<const uint4*>(w + i));
        const uint32_t a[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w};
        uint4 o;
        uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e)
            op[e] = pack_bf16(bf16_lo(a[e]) * inv * bf16_lo(gg[e]), bf16_hi(a[e]) * inv * bf16_hi(gg[e]));
        *reinterpret_cast<uint4*>(xs + i) = o;
    }
}
__device__ __forceinline__ void load_x_copy(const __nv_bfloat16* src, __nv_bfloat16* xs, int n) {
    for (int i = threadIdx.x * 8; i < n; i += GV_CONSUMERS * 8)
        *reinterpret_cast<uint4*>(xs + i) = *reinterpret_cast<const uint4*>(src + i);
}

struct FinishState {
    float best;
    int best_idx;
};

// Shared-memory context handed to the ops.
struct GvSmem {
    __nv_bfloat16* xs;   // activation vector(s)
    float* scratch;      // 64 floats
    float* prep;         // GV_PREP_FLOATS floats of per-CTA epilogue constants
    int* misc;           // 8 ints (e.g. selected expert ids, position, slot)
    uint64_t* aux_bar;   // producer gate for ops whose row addresses depend on the prologue (router)
};

template <class Op>
__global__ void __launch_bounds__(GV_THREADS, 1)
stream_gemv_kernel(const Op op) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(ring + GV_STAGES * GV_STAGE_BYTES);
    uint8_t* tail = reinterpret_cast<uint8_t*>(xs) + ((op.x_elems() * 2 + 127) / 128) * 128;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + GV_STAGES;
    uint64_t* aux_bar = empty_bar + GV_STAGES;
    float* red = reinterpret_cast<float*>(aux_bar + 2);   // [2 groups][2 parities][8 warps][8]
    float* scratch = red + 2 * 2 * 8 * 8;                 // 64
    float* prep = scratch + 64;                           // GV_PREP_FLOATS
    int* misc = reinterpret_cast<int*>(prep + GV_PREP_FLOATS);
    GvSmem sm{xs, scratch, prep, misc, aux_bar};

    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = op.K;
    const int n_chunks = (K + GV_KC - 1) / GV_KC;
    const long long n_items = op.num_items();
    const int i0 = static_cast<int>(n_items * blockIdx.x / gridDim.x);
    const int i1 = static_cast<int>(n_items * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int i = 0; i < GV_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], GV_GROUP_WARPS);
        }
        mbar_init(aux_bar, 1);
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == GV_CONSUMERS / 32) {
        if (lane == 0) {
            if (Op::kProducerReadsUpstream) pdl_wait();   // row addresses depend on the previous kernel's output
            if (Op::kProducerNeedsPrologue) mbar_wait(aux_bar, 0, 13);
            int stage = 0;
            uint32_t phase = 0;
            for (int item = i0; item < i1; ++item) {
                const __nv_bfloat16* r0 = op.row_ptr(b, item, 0, sm);
                const __nv_bfloat16* r1 = op.row_ptr(b, item, 1, sm);
                for (int c = 0; c < n_chunks; ++c) {
                    const int len = min(GV_KC, K - c * GV_KC);
                    mbar_wait(&empty_bar[stage], phase ^ 1, 11);
                    mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(len) * 4);
                    uint8_t* dst = ring + stage * GV_STAGE_BYTES;
                    bulk_copy_g2s(dst, r0 + c * GV_KC, static_cast<uint32_t>(len) * 2, &full_bar[stage]);
                    bulk_copy_g2s(dst + GV_KC * 2, r1 + c * GV_KC, static_cast<uint32_t>(len) * 2, &full_bar[stage]);
                    if (++stage == GV_STAGES) { stage = 0; phase ^= 1; }
                }
            }
            // every weight load of this CTA is issued: the next kernel of the chain may start filling its ring.
            // Waiting first keeps the chain transitive (when kernel N+1 starts, kernel N-1 is complete).
            pdl_wait();
            pdl_launch_dependents();
        }
        return;
    }

    // ---------------------------------------------------------------- consumers (512 threads, 2 groups of 8 warps)
    pdl_wait();   // activations come from the previous kernel; weights (producer warp) do not
    op.prologue(b, sm, i0, i1);
    consumer_barrier_all();
    FinishState st{-INFINITY, 0x7fffffff};
    const int group = warp >> 3, gw = warp & 7;
    const int NB = (n_chunks == 1) ? GV_NB : 1;
    int stage = 0;
    uint32_t phase = 0;
    int bi = 0;
    for (int item0 = i0; item0 < i1; item0 += NB, ++bi) {
        const int nb = min(NB, i1 - item0);
        if ((bi & 1) != group) {  // the other group's batch: skip its stages
            stage += nb * n_chunks;
            while (stage >= GV_STAGES) { stage -= GV_STAGES; phase ^= 1; }
            continue;
        }
        float acc[GV_NB][2];
#pragma unroll
        for (int q = 0; q < GV_NB; ++q) { acc[q][0] = 0.0f; acc[q][1] = 0.0f; }
#pragma unroll
        for (int q = 0; q < GV_NB; ++q) {
            if (q < nb) {
                for (int c = 0; c < n_chunks; ++c) {
                    const int len = min(GV_KC, K - c * GV_KC);
                    mbar_wait(&full_bar[stage], phase, 12);
                    const uint8_t* w0p = ring + stage * GV_STAGE_BYTES;
                    const uint8_t* w1p = w0p + GV_KC * 2;
#pragma unroll
                    for (int pss = 0; pss < 2; ++pss) {
                        const int off = gw * 512 + pss * 256 + lane * 8;
                        if (off < len) {
                            const uint4 w0 = *reinterpret_cast<const uint4*>(w0p + off * 2);
                            const uint4 w1 = *reinterpret_cast<const uint4*>(w1p + off * 2);
                            const uint4 x0 = *reinterpret_cast<const uint4*>(xs + c * GV_KC + off);
                            const uint32_t a[4] = {w0.x, w0.y, w0.z, w0.w}, bb[4] = {w1.x, w1.y, w1.z, w1.w};
                            const uint32_t xa[4] = {x0.x, x0.y, x0.z, x0.w};
                            if constexpr (Op::kXPerRow) {
                                const uint4 x1 = *reinterpret_cast<const uint4*>(xs + K + c * GV_KC + off);
                                const uint32_t xb[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    acc[q][0] += bf16_lo(a[e]) * bf16_lo(xa[e]) + bf16_hi(a[e]) * bf16_hi(xa[e]);
                                    acc[q][1] += bf16_lo(bb[e]) * bf16_lo(xb[e]) + bf16_hi(bb[e]) * bf16_hi(xb[e]);
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float xl = bf16_lo(xa[e]), xh = bf16_hi(xa[e]);
                                    acc[q][0] += bf16_lo(a[e]) * xl + bf16_hi(a[e]) * xh;
                                    acc[q][1] += bf16_lo(bb[e]) * xl + bf16_hi(bb[e]) * xh;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[stage]);
                    if (++stage == GV_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < GV_NB; ++q) {
            acc[q][0] = warp_sum(acc[q][0]);
            acc[q][1] = warp_sum(acc[q][1]);
        }
        float* rb = red + ((group * 2 + ((bi >> 1) & 1)) * 8) * 8;
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < GV_NB; ++q) { rb[gw * 8 + q * 2] = acc[q][0]; rb[gw * 8 + q * 2 + 1] = acc[q][1]; }
        }
        group_barrier(group);
        if (gw < nb && lane == 0) {
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { s0 += rb[w * 8 + gw * 2]; s1 += rb[w * 8 + gw * 2 + 1]; }
            op.finish(b, item0 + gw, s0, s1, st, sm, i0);
        }
    }
    op.finalize(b, st);
}

// ------------------------------------------------------------------------------------------------ ops
struct QkvOp {
    static constexpr bool kProducerReadsUpstream = false;
    static constexpr bool kXPerRow = false;
    static constexpr bool kProducerNeedsPrologue = false;
    const __nv_bfloat16* h;       // [B, H]
    const __nv_bfloat16* norm_w;  // [H]
    const __nv_bfloat16* w_qkv;   // [(n_q + 2 n_kv) * 128, H]
    const float* cos_sin;         // [max_pos, 2, 64]
    const int* cur_pos;           // [B]
    const int* block_table;       // [B, max_pages]
    __nv_bfloat16* q_out;         // [B, n_q * 128]
    __nv_bfloat16* k_cache;       // [slots, n_kv, 128]
    __nv_bfloat16* v_cache;
    int K, n_q, n_kv, page_size, max_pages;
    float eps;

    __device__ int x_elems() const { return K; }
    __device__ long long num_items() const { return static_cast<long long>(n_q + 2 * n_kv) * 64; }
    __device__ const __nv_bfloat16* row_ptr(int, int item, int r, const GvSmem&) const {
        const int head = item >> 6, j = item & 63;
        return w_qkv + static_cast<long long>(head * 128 + j + r * 64) * K;
    }
    __device__ void prologue(int b, const GvSmem& sm, int, int) const {
        // epilogue constants: position, cache slot, the cos/sin row of this position
        const int pos = cur_pos[b];
        if (threadIdx.x < 128) sm.prep[threadIdx.x] = cos_sin[static_cast<long long>(pos) * 128 + threadIdx.x];
        if (threadIdx.x == 0) {
            const int page = block_table[static_cast<long long>(b) * max_pages + pos / page_size];
            sm.misc[0] = page * page_size + pos % page_size;
        }
        load_x_rmsnorm(h + static_cast<long long>(b) * K, norm_w, sm.xs, K, eps, sm.scratch);
    }
    __device__ void finish(int b, int item, float s0, float s1, FinishState&, const GvSmem& sm, int) const {
        const int head = item >> 6, j = item & 63;
        // qkv projections are rounded to bf16 before RoPE, as the GEMM path (and the reference) does
        s0 = __bfloat162float(__float2bfloat16(s0));
        s1 = __bfloat162float(__float2bfloat16(s1));
        if (head < n_q + n_kv) {
            const float c = sm.prep[j], s = sm.prep[64 + j];
            const float o0 = s0 * c - s1 * s, o1 = s1 * c + s0 * s;
            s0 = o0;
            s1 = o1;
        }
        if (head < n_q) {
            __nv_bfloat16* q = q_out + (static_cast<long long>(b) * n_q + head) * 128;
            q[j] = __float2bfloat16(s0);
            q[j + 64] = __float2bfloat16(s1);
        } else {
            const long long slot = sm.misc[0];
            const bool is_k = head < n_q + n_kv;
            const int kvh = is_k ? head - n_q : head - n_q - n_kv;
            __nv_bfloat16* dst = (is_k ? k_cache : v_cache) + (slot * n_kv + kvh) * 128;
            dst[j] = __float2bfloat16(s0);
            dst[j + 64] = __float2bfloat16(s1);
        }
    }
    __device__ void finalize(int, FinishState&) const {}
};

struct OProjOp {
    static constexpr bool kProducerReadsUpstream = false;
    static constexpr bool kXPerRow = false;
    static constexpr bool kProducerNeedsPrologue = false;
    const __nv_bfloat16* x;  // [B, K] attention output
    const __nv_bfloat16* w;  // [N, K]
    __nv_bfloat16* h;        // [B, N] residual stream, updated in place
    int K, N;

    __device__ int x_elems() const { return K; }
    __device__ long long num_items() const { return N / 2; }
    __device__ const __nv_bfloat16* row_ptr(int, int item, int r, const GvSmem&) const {
        return w + static_cast<long long>(item * 2 + r) * K;
    }
    __device__ void prologue(int b, const GvSmem& sm, int i0, int i1) const {
        const int n = (i1 - i0) * 2;   // residual values of this CTA's rows, prefetched so finish() never waits
        if (n <= GV_PREP_FLOATS)
            for (int i = threadIdx.x; i < n; i += GV_CONSUMERS)
                sm.prep[i] = __bfloat162float(h[static_cast<long long>(b) * N + i0 * 2 + i]);
        load_x_copy(x + static_cast<long long>(b) * K, sm.xs, K);
    }
    __device__ void finish(int b, int item, float s0, float s1, FinishState&, const GvSmem& sm, int i0) const {
        __nv_bfloat16* hr = h + static_cast<long long>(b) * N + item * 2;
        float r0, r1;
        if ((static_cast<int>(num_items() * (blockIdx.x + 1) / gridDim.x) - i0) * 2 <= GV_PREP_FLOATS) {
            r0 = sm.prep[(item - i0) * 2];
            r1 = sm.prep[(item - i0) * 2 + 1];
        } else {
            r0 = __bfloat162float(hr[0]);
            r1 = __bfloat162float(hr[1]);
        }
        *reinterpret_cast<uint32_t*>(hr) = pack_bf16(r0 + s0, r1 + s1);
    }
    __device__ void finalize(int, FinishState&) const {}
};

// post_attention_layernorm + router (top-2 of 8, fp32 softmax, renormalised) fused into the expert gate/up GEMV:
// every CTA recomputes the 8 router logits (64 KB of L2-resident gate weights) instead of paying a kernel boundary.
struct GateUpOp {
    static constexpr bool kProducerReadsUpstream = false;
    static constexpr bool kXPerRow = false;
    static constexpr bool kProducerNeedsPrologue = true;
    const __nv_bfloat16* h;        // [B, H] residual stream (post attention)
    const __nv_bfloat16* norm_w;   // [H]
    const __nv_bfloat16* gate_w;   // [8, H]
    const __nv_bfloat16* w13;      // [E, 2I, H]
    int* topk_ids;                 // [B, 2]  (written by CTA 0 for the down kernel)
    float* topk_w;                 // [B, 2]
    __nv_bfloat16* act;            // [B, 2, I]
    int K, I;
    float eps;

    __device__ int x_elems() const { return K; }
    __device__ long long num_items() const { return 2ll * I; }
    __device__ const __nv_bfloat16* row_ptr(int, int item, int r, const GvSmem& sm) const {
        const int k = item / I, j = item % I;
        const int e = sm.misc[k];
        return w13 + (static_cast<long long>(e) * 2 * I + r * I + j) * K;
    }
    __device__ void prologue(int b, const GvSmem& sm, int, int) const {
        const __nv_bfloat16* hr = h + static_cast<long long>(b) * K;
        float part[9];   // sum of squares + 8 un-normalised logits  sum_i (h_i * w_i) * g_ei
#pragma unroll
        for (int e = 0; e < 9; ++e) part[e] = 0.0f;
        for (int i = threadIdx.x * 8; i < K; i += GV_CONSUMERS * 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(hr + i);
            const uint4 g = __ldg(reinterpret_cast<const uint4*>(norm_w + i));
            uint4 ge[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) ge[e] = __ldg(reinterpret_cast<const uint4*>(gate_w + static_cast<long long>(e) * K + i));
            const uint32_t a[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w};
            float xw[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float lo = bf16_lo(a[q]), hi = bf16_hi(a[q]);
                part[0] += lo * lo + hi * hi;
                xw[2 * q] = lo * bf16_lo(gg[q]);
                xw[2 * q + 1] = hi * bf16_hi(gg[q]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t w[4] = {ge[e].x, ge[e].y, ge[e].z, ge[e].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) part[1 + e] += xw[2 * q] * bf16_lo(w[q]) + xw[2 * q + 1] * bf16_hi(w[q]);
            }
        }
#pragma unroll
        for (int e = 0; e < 9; ++e) part[e] = warp_sum(part[e]);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float* red9 = sm.prep;   // [16 warps][9]
        if (lane == 0)
#pragma unroll
            for (int e = 0; e < 9; ++e) red9[warp * 9 + e] = part[e];
        consumer_barrier_all();
        float tot[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) {
            tot[e] = 0.0f;
            for (int w = 0; w < 16; ++w) tot[e] += red9[w * 9 + e];
        }
        const float inv = rsqrtf(tot[0] / static_cast<float>(K) + eps);
        float p[8], m = -INFINITY, sum = 0.0f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { p[e] = tot[1 + e] * inv; m = fmaxf(m, p[e]); }
#pragma unroll
        for (int e = 0; e < 8; ++e) { p[e] = expf(p[e] - m); sum += p[e]; }
        int e0 = 0;
#pragma unroll
        for (int e = 1; e < 8; ++e) if (p[e] > p[e0]) e0 = e;
        int e1 = (e0 == 0) ? 1 : 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) if (e != e0 && p[e] > p[e1]) e1 = e;
        if (threadIdx.x == 0) {
            sm.misc[0] = e0;
            sm.misc[1] = e1;
            mbar_arrive(sm.aux_bar);   // release the producer: the rows to stream are now known
            if (blockIdx.x == 0) {
                const float p0 = p[e0] / sum, p1 = p[e1] / sum, den = p0 + p1;
                topk_ids[b * 2] = e0;
                topk_ids[b * 2 + 1] = e1;
                topk_w[b * 2] = p0 / den;
                topk_w[b * 2 + 1] = p1 / den;
            }
        }
        // normalised activations (bf16) for the GEMV
        for (int i = threadIdx.x * 8; i < K; i += GV_CONSUMERS * 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(hr + i);
            const uint4 g = __ldg(reinterpret_cast<const uint4*>(norm_w + i));
            const uint32_t a[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w};
            uint4 o;
            uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                op[q] = pack_bf16(bf16_lo(a[q]) * inv * bf16_lo(gg[q]), bf16_hi(a[q]) * inv * bf16_hi(gg[q]));
            *reinterpret_cast<uint4*>(sm.xs + i) = o;
        }
    }
    __device__ void finish(int b, int item, float s0, float s1, FinishState&, const GvSmem&, int) const {
        // gate and up are bf16 linear outputs in the reference; silu(gate) * up in fp32, one rounding
        act[static_cast<long long>(b) * 2 * I + item] = __float2bfloat16(silu(s0) * s1);
    }
    __device__ void finalize(int, FinishState&) const {}
};

struct DownOp {
    static constexpr bool kProducerReadsUpstream = true;
    static constexpr bool kXPerRow = true;
    static constexpr bool kProducerNeedsPrologue = false;
    const __nv_bfloat16* act;   // [B, 2, I]
    const __nv_bfloat16* w2;    // [E, H, I]
    const int* topk_ids;        // [B, 2]
    const float* topk_w;        // [B, 2]
    __nv_bfloat16* h;           // [B, H] residual stream, updated in place
    int K, H;                   // K = I

    __device__ int x_elems() const { return 2 * K; }
    __device__ long long num_items() const { return H; }
    __device__ const __nv_bfloat16* row_ptr(int b, int item, int r, const GvSmem&) const {
        const int e = topk_ids[b * 2 + r];
        return w2 + (static_cast<long long>(e) * H + item) * K;
    }
    __device__ void prologue(int b, const GvSmem& sm, int i0, int i1) const {
        const int n = i1 - i0;
        if (n <= GV_PREP_FLOATS - 2)
            for (int i = threadIdx.x; i < n; i += GV_CONSUMERS)
                sm.prep[i] = __bfloat162float(h[static_cast<long long>(b) * H + i0 + i]);
        if (threadIdx.x < 2) sm.prep[GV_PREP_FLOATS - 2 + threadIdx.x] = topk_w[b * 2 + threadIdx.x];
        load_x_copy(act + static_cast<long long>(b) * 2 * K, sm.xs, 2 * K);
    }
    __device__ void finish(int b, int item, float s0, float s1, FinishState&, const GvSmem& sm, int i0) const {
        __nv_bfloat16* hr = h + static_cast<long long>(b) * H + item;
        const int n = static_cast<int>(num_items() * (blockIdx.x + 1) / gridDim.x) - i0;
        const float r = (n <= GV_PREP_FLOATS - 2) ? sm.prep[item - i0] : __bfloat162float(hr[0]);
        const float y = sm.prep[GV_PREP_FLOATS - 2] * s0 + sm.prep[GV_PREP_FLOATS - 1] * s1;
        hr[0] = __float2bfloat16(r + y);
    }
    __device__ void finalize(int, FinishState&) const {}
};

__device__ __forceinline__ unsigned long long pack_argmax(float v, int idx) {
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(idx));
}

struct LmHeadOp {
    static constexpr bool kProducerReadsUpstream = false;
    static constexpr bool kXPerRow = false;
    static constexpr bool kProducerNeedsPrologue = false;
    const __nv_bfloat16* h;        // rows of the residual stream, row b at h + b * h_stride
    long long h_stride;
    const __nv_bfloat16* norm_w;   // final RMSNorm
    const __nv_bfloat16* w;        // [V, K]
    __nv_bfloat16* logits;         // [B, V] or nullptr
    unsigned long long* best;      // [B] packed (value, ~index), reset to 0 before the step
    int K, V;
    float eps;

    __device__ int x_elems() const { return K; }
    __device__ long long num_items() const { return (V + 1) / 2; }
    __device__ const __nv_bfloat16* row_ptr(int, int item, int r, const GvSmem&) const {
        int row = item * 2 + r;
        if (row >= V) row = V - 1;
        return w + static_cast<long long>(row) * K;
    }
    __device__ void prologue(int b, const GvSmem& sm, int, int) const {
        load_x_rmsnorm(h + static_cast<long long>(b) * h_stride, norm_w, sm.xs, K, eps, sm.scratch);
    }
    __device__ void finish(int b, int item, float s0, float s1, FinishState& st, const GvSmem&, int) const {
        const int r0 = item * 2, r1 = item * 2 + 1;
        // logits stay in the activation dtype and arg-max runs on them (vita_mixtral.py:171-173)
        const __nv_bfloat16 l0 = __float2bfloat16(s0), l1 = __float2bfloat16(s1);
        if (logits) {
            logits[static_cast<long long>(b) * V + r0] = l0;
            if (r1 < V) logits[static_cast<long long>(b) * V + r1] = l1;
        }
        const float f0 = __bfloat162float(l0), f1 = __bfloat162float(l1);
        if (f0 > st.best || (f0 == st.best && r0 < st.best_idx)) { st.best = f0; st.best_idx = r0; }
        if (r1 < V && (f1 > st.best || (f1 == st.best && r1 < st.best_idx))) { st.best = f1; st.best_idx = r1; }
    }
    __device__ void finalize(int b, FinishState& st) const {
        if ((threadIdx.x & 31) == 0 && st.best_idx != 0x7fffffff) atomicMax(&best[b], pack_argmax(st.best, st.best_idx));
    }
};

template <class Op>
static int launch_stream_gemv(const Op& op, int x_elems, int B, cudaStream_t st, const char* name) {
    const int smem_bytes = GV_STAGES * GV_STAGE_BYTES + ((x_elems * 2 + 127) / 128) * 128 + (2 * GV_STAGES + 2) * 8 +
                           (2 * 2 * 8 * 8 + 64 + GV_PREP_FLOATS + 8) * 4 + 256;
    auto kern = stream_gemv_kernel<Op>;
    static int configured_bytes = 0;
    if (smem_bytes > configured_bytes) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes), name);
        if (rc) return rc;
        configured_bytes = smem_bytes;
    }
    dim3 grid(num_sms(), B);
    cudaError_t e = launch_chain(kern, grid, dim3(GV_THREADS), smem_bytes, st, op);
    if (e != cudaSuccess) return check_cuda(e, name);
    return check_launch(name);
}

// ------------------------------------------------------------------------------------------------ small kernels
// Start of a decode step: consume the previous arg-max, log it, advance the cache length, gather the embedding.
__global__ void __launch_bounds__(256)
decode_embed_kernel(unsigned long long* best, int* token_log, int* gen_count, int max_log, int* cache_len,
                    int* cur_pos, const __nv_bfloat16* embed, __nv_bfloat16* h, int H, int vocab, int max_ctx,
                    unsigned long long* chain_serial) {
    const int b = blockIdx.x;
    __shared__ int s_tok;
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0) {
        if (b == 0 && chain_serial != nullptr) *chain_serial += 1;   // the step's kernels poll serial x arrivals
        const unsigned long long key = best[b];
        int tok = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull));
        if (tok < 0 || tok >= vocab) tok = 0;
        s_tok = tok;
        const int n = gen_count[b];
        if (n < max_log) token_log[static_cast<long long>(b) * max_log + n] = tok;
        gen_count[b] = n + 1;
        // a full cache stays full: the position saturates on the last slot instead of running into the next
        // sequence's block-table row / beyond the rope table (the host refuses such requests up front)
        const int len = cache_len[b];
        cur_pos[b] = len < max_ctx ? len : max_ctx - 1;
        cache_len[b] = len < max_ctx ? len + 1 : max_ctx;
        best[b] = 0ull;
    }
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(embed + static_cast<long long>(s_tok) * H);
    uint4* dst = reinterpret_cast<uint4*>(h + static_cast<long long>(b) * H);
    for (int i = threadIdx.x; i < (H >> 3); i += blockDim.x) dst[i] = src[i];
}

// post_attention_layernorm + router for one token per block (E = 8 warps).
__global__ void __launch_bounds__(256)
decode_router_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ norm_w,
                     const __nv_bfloat16* __restrict__ gate_w, __nv_bfloat16* __restrict__ xn,
                     int* __restrict__ topk_ids, float* __restrict__ topk_w, int H, float eps) {
    extern __shared__ __align__(16) uint8_t smem_r[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_r);
    __shared__ float red[8];
    __shared__ float logits[8];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __nv_bfloat16* hr = h + static_cast<long long>(b) * H;
    float ss = 0.0f;
    for (int i = threadIdx.x; i < H; i += 256) { const float v = __bfloat162float(hr[i]); ss += v * v; }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float inv = rsqrtf(tot / static_cast<float>(H) + eps);
    for (int i = threadIdx.x; i < H; i += 256) {
        const __nv_bfloat16 v = __float2bfloat16(__bfloat162float(hr[i]) * inv * __bfloat162float(norm_w[i]));
        xs[i] = v;
        xn[static_cast<long long>(b) * H + i] = v;
    }
    __syncthreads();
    float acc = 0.0f;
    const __nv_bfloat16* gw = gate_w + static_cast<long long>(warp) * H;
    for (int i = lane * 8; i < H; i += 256) {
        const uint4 w = *reinterpret_cast<const uint4*>(gw + i);
        const uint4 x = *reinterpret_cast<const uint4*>(xs + i);
        const uint32_t wa[4] = {w.x, w.y, w.z, w.w}, xa[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) acc += bf16_lo(wa[e]) * bf16_lo(xa[e]) + bf16_hi(wa[e]) * bf16_hi(xa[e]);
    }
    acc = warp_sum(acc);
    if (lane == 0) logits[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = -INFINITY;
        for (int e = 0; e < 8; ++e) m = fmaxf(m, logits[e]);
        float p[8], sum = 0.0f;
        for (int e = 0; e < 8; ++e) { p[e] = expf(logits[e] - m); sum += p[e]; }
        int i0 = 0;
        for (int e = 1; e < 8; ++e) if (p[e] > p[i0]) i0 = e;
        int i1 = (i0 == 0) ? 1 : 0;
        for (int e = 0; e < 8; ++e) if (e != i0 && p[e] > p[i1]) i1 = e;
        const float p0 = p[i0] / sum, p1 = p[i1] / sum, den = p0 + p1;
        topk_ids[b * 2] = i0;
        topk_ids[b * 2 + 1] = i1;
        topk_w[b * 2] = p0 / den;
        topk_w[b * 2 + 1] = p1 / den;
    }
}

// Row-wise arg-max of bf16 logits (first index on ties, like torch.argmax), packed like the GEMV kernels do.
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const __nv_bfloat16* __restrict__ logits, unsigned long long* __restrict__ best, int V) {
    __shared__ unsigned long long red[8];
    const __nv_bfloat16* row = logits + static_cast<long long>(blockIdx.x) * V;
    unsigned long long loc = 0ull;
    for (int i = threadIdx.x; i < V; i += 256) {
        const unsigned long long k = pack_argmax(__bfloat162float(row[i]), i);
        loc = k > loc ? k : loc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, loc, o);
        loc = other > loc ? other : loc;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = red[0];
        for (int i = 1; i < 8; ++i) m = red[i] > m ? red[i] : m;
        best[blockIdx.x] = m;
    }
}

// slots[b] = paged-KV slot of position cur_pos[b] of sequence b (batched decode through the GEMM path)
__global__ void decode_slots_kernel(const int* __restrict__ cur_pos, const int* __restrict__ block_table,
                                    int* __restrict__ slots, int B, int page_size, int max_pages) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int pos = cur_pos[b];
    slots[b] = block_table[static_cast<long long>(b) * max_pages + pos / page_size] * page_size + pos % page_size;
}

// L2 prefetch of a weight matrix (cp.async.bulk.prefetch.L2): lets the o-projection weights arrive while the
// (latency-bound) decode attention runs, so the o-projection GEMV then streams from L2.
__global__ void l2_prefetch_kernel(const char* __restrict__ base, long long bytes) {
    const long long chunk = 64 * 1024;
    const long long n_chunks = (bytes + chunk - 1) / chunk;
    for (long long c = blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long off = c * chunk;
        const unsigned int sz = static_cast<unsigned int>((bytes - off < chunk ? bytes - off : chunk) & ~15ll);
        if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(sz) : "memory");
    }
}

}  // namespace vita

using namespace vita;

extern "C" int vita_l2_prefetch(const void* ptr, int64_t bytes, void* stream) {
    VITA_REQUIRE(aligned16(ptr), "pointer must be 16-byte aligned");
    if (bytes <= 0) return VITA_OK;
    l2_prefetch_kernel<<<num_sms() > 0 ? num_sms() : 1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const char*>(ptr), bytes);
    return check_launch("l2_prefetch");
}

extern "C" int vita_argmax_rows(const void* logits, uint64_t* best, int64_t B, int64_t V, void* stream) {
    if (B == 0) return VITA_OK;
    argmax_rows_kernel<<<static_cast<unsigned>(B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        BF16C(logits), reinterpret_cast<unsigned long long*>(best), (int)V);
    return check_launch("argmax_rows");
}

extern "C" int vita_decode_slots(const int32_t* cur_pos, const int32_t* block_table, int32_t* slots, int64_t B,
                                 int64_t page_size, int64_t max_pages, void* stream) {
    if (B == 0) return VITA_OK;
    decode_slots_kernel<<<static_cast<unsigned>((B + 63) / 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(
        cur_pos, block_table, slots, (int)B, (int)page_size, (int)max_pages);
    return check_launch("decode_slots");
}

extern "C" int vita_decode_embed(uint64_t* best, int32_t* token_log, int32_t* gen_count, int64_t max_log,
                                 int32_t* cache_len, int32_t* cur_pos, const void* embed, void* h, int64_t B,
                                 int64_t H, int64_t vocab, int64_t max_ctx, uint64_t* chain_serial, void* stream) {
    VITA_REQUIRE(H % 8 == 0, "H must be a multiple of 8");
    VITA_REQUIRE(max_ctx > 0 && max_ctx <= 0x7fffffff, "max_ctx (KV capacity per sequence) must be positive");
    if (B == 0) return VITA_OK;
    cudaError_t e = launch_chain(decode_embed_kernel, dim3(static_cast<unsigned>(B)), dim3(256), 0,
                                 static_cast<cudaStream_t>(stream), reinterpret_cast<unsigned long long*>(best),
                                 token_log, gen_count, (int)max_log, cache_len, cur_pos, BF16C(embed),
                                 static_cast<__nv_bfloat16*>(h), (int)H, (int)vocab, (int)max_ctx,
                                 reinterpret_cast<unsigned long long*>(chain_serial));
    if (e != cudaSuccess) return check_cuda(e, "decode_embed");
    return check_launch("decode_embed");
}

extern "C" int vita_decode_qkv_rope(const void* h, const void* norm_w, const void* w_qkv, const float* cos_sin,
                                    const int32_t* cur_pos, const int32_t* block_table, void* q_out, void* k_cache,
                                    void* v_cache, int64_t B, int64_t H, int64_t n_q_heads, int64_t n_kv_heads,
                                    int64_t head_dim, int64_t page_size, int64_t max_pages, float eps, void* stream) {
    VITA_REQUIRE(head_dim == 128, "head_dim must be 128");
    VITA_REQUIRE(H % 8 == 0, "H must be a multiple of 8");
    if (B == 0) return VITA_OK;
    QkvOp op{BF16C(h), BF16C(norm_w), BF16C(w_qkv), cos_sin, cur_pos, block_table, static_cast<__nv_bfloat16*>(q_out),
             static_cast<__nv_bfloat16*>(k_cache), static_cast<__nv_bfloat16*>(v_cache), (int)H, (int)n_q_heads,
             (int)n_kv_heads, (int)page_size, (int)max_pages, eps};
    return launch_stream_gemv(op, (int)H, (int)B, static_cast<cudaStream_t>(stream), "decode_qkv_rope");
}

extern "C" int vita_decode_oproj(const void* x, const void* w, void* h, int64_t B, int64_t N, int64_t K,
                                 void* stream) {
    VITA_REQUIRE(K % 8 == 0 && N % 2 == 0, "K must be a multiple of 8 and N even");
    if (B == 0) return VITA_OK;
    OProjOp op{BF16C(x), BF16C(w), static_cast<__nv_bfloat16*>(h), (int)K, (int)N};
    return launch_stream_gemv(op, (int)K, (int)B, static_cast<cudaStream_t>(stream), "decode_oproj");
}

extern "C" int vita_decode_router(const void* h, const void* norm_w, const void* gate_w, void* xn, int32_t* topk_ids,
                                  float* topk_w, int64_t B, int64_t H, int64_t E, float eps, void* stream) {
    VITA_REQUIRE(E == 8, "router is specialised for 8 experts (Mixtral-8x7B)");
    VITA_REQUIRE(H % 8 == 0 && H * 2 <= 48 * 1024, "H must be a multiple of 8 and fit 48 KB of shared memory");
    if (B == 0) return VITA_OK;
    decode_router_kernel<<<static_cast<unsigned>(B), 256, H * 2, static_cast<cudaStream_t>(stream)>>>(
        BF16C(h), BF16C(norm_w), BF16C(gate_w), static_cast<__nv_bfloat16*>(xn), topk_ids, topk_w, (int)H, eps);
    return check_launch("decode_router");
}

extern "C" int vita_decode_moe_gate_up(const void* h, const void* norm_w, const void* gate_w, const void* w13,
                                       int32_t* topk_ids, float* topk_w, void* act, int64_t B, int64_t H, int64_t I,
                                       int64_t E, float eps, void* stream) {
    VITA_REQUIRE(H % 8 == 0, "H must be a multiple of 8");
    VITA_REQUIRE(E == 8, "router is specialised for 8 experts (Mixtral-8x7B)");
    if (B == 0) return VITA_OK;
    GateUpOp op{BF16C(h), BF16C(norm_w), BF16C(gate_w), BF16C(w13), topk_ids, topk_w,
                static_cast<__nv_bfloat16*>(act), (int)H, (int)I, eps};
    return launch_stream_gemv(op, (int)H, (int)B, static_cast<cudaStream_t>(stream), "decode_moe_gate_up");
}

extern "C" int vita_decode_moe_down(const void* act, const void* w2, const int32_t* topk_ids, const float* topk_w,
                                    void* h, int64_t B, int64_t H, int64_t I, void* stream) {
    VITA_REQUIRE(I % 8 == 0, "I must be a multiple of 8");
    if (B == 0) return VITA_OK;
    DownOp op{BF16C(act), BF16C(w2), topk_ids, topk_w, static_cast<__nv_bfloat16*>(h), (int)I, (int)H};
    return launch_stream_gemv(op, (int)(2 * I), (int)B, static_cast<cudaStream_t>(stream), "decode_moe_down");
}

extern "C" int vita_lm_head_argmax(const void* h, int64_t h_stride, const void* norm_w, const void* w, void* logits,
                                   uint64_t* best, int64_t B, int64_t H, int64_t V, float eps, void* stream) {
    VITA_REQUIRE(H % 8 == 0, "H must be a multiple of 8");
    if (B == 0) return VITA_OK;
    LmHeadOp op{BF16C(h), h_stride, BF16C(norm_w), BF16C(w), static_cast<__nv_bfloat16*>(logits),
                reinterpret_cast<unsigned long long*>(best), (int)H, (int)V, eps};
    return launch_stream_gemv(op, (int)H, (int)B, static_cast<cudaStream_t>(stream), "lm_head_argmax");
}
