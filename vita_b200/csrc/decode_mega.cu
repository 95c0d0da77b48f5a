// One persistent kernel per decode step (bs = 1): every linear of all layers, the paged attention and the LM head
// run as PHASES of a single launch, separated by grid barriers instead of kernel boundaries.
//
// Why: with one kernel per linear (decode_tc.cu) each of the 5 kernels of a layer pays ~9 us of launch gap, barrier /
// TMEM setup, prologue, pipeline fill and stream-K tail -- ~45 us per layer on top of the 120 us the weights take to
// stream (ncu launch list, profiles/).  Here the TMA producer warp never stops: weight tiles of the NEXT phase are
// already in the shared-memory ring while the epilogue warps finish the current phase, cross the grid barrier and
// compute the next activation vector.  The only data dependency the producer has is on the two expert ids of the
// layer (known after the router, which every CTA recomputes in the gate/up prologue).
//
//   warp 0   TMA producer (weights; uniform ring stage = 2 weight tiles of 128 rows x 64 k + 2 activation tiles)
//   warp 1   MMA issuer   (tcgen05.mma M=128 N=16 K=16, accumulators double-buffered in TMEM)
//   warp 2   TMEM allocator
//   warp 3   x-tile writer (row 0 of the 16 x 64 N-operand tiles)
//   warps 4-7 per phase: grid barrier -> prologue (RMSNorm / router / copy) -> per-segment epilogues (RoPE + KV append,
//            residual, SiLU*up, weighted combine, logits + arg-max) and the whole attention phase.
// Phase order per layer: QKV -> ATTN -> OPROJ -> GATEUP -> DOWN ; after the last layer: LMHEAD.
// Work split: stream-K over (row block, k block) units, deterministic partial slots + tickets (as decode_tc.cu).
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace vita {

constexpr int MG_THREADS = 256;
constexpr int MG_STAGES = 5;
constexpr int MG_A_BYTES = 128 * 64 * 2;
constexpr int MG_X_BYTES = 16 * 64 * 2;
constexpr int MG_STAGE_A = 2 * MG_A_BYTES;
constexpr int MG_STAGE_X = 2 * MG_X_BYTES;
constexpr int MG_SLOTS = 8;
constexpr int MG_D = 128;      // head dim
constexpr int MG_GROUP = 4;    // q heads per kv head
constexpr int MG_MAXS = 16;    // attention context splits

enum { PH_QKV = 0, PH_ATTN = 1, PH_OPROJ = 2, PH_GATEUP = 3, PH_DOWN = 4, PH_LMHEAD = 5 };
enum { MODE_SEQ = 0, MODE_DUAL_SAMEX = 1, MODE_DUAL = 2 };

struct MegaLayer {
    const __nv_bfloat16* ln1;
    const __nv_bfloat16* ln2;
    const __nv_bfloat16* gate;
    __nv_bfloat16* k_cache;
    __nv_bfloat16* v_cache;
};

struct MegaParams {
    const CUtensorMap* maps;   // [n_layers * 4 + 1]: qkv, o, w13, w2 per layer, then lm_head
    const MegaLayer* layers;
    int n_layers;
    const __nv_bfloat16* final_norm;
    int H, I, n_q, n_kv, V, page_size, max_pages, splits;
    float eps, attn_scale_log2;
    __nv_bfloat16* h;          // [H] residual stream
    __nv_bfloat16* q;          // [n_q * 128]
    __nv_bfloat16* attn;       // [n_q * 128]
    __nv_bfloat16* act;        // [2, I]
    __nv_bfloat16* logits;     // [V] or nullptr
    unsigned long long* best;
    const float* cos_sin;
    const int* cur_pos;
    const int* block_table;
    float* scratch;            // [max_rb * SLOTS * 2 * 128]
    int* tickets;              // [max_rb]
    unsigned int* grid_bar;    // zeroed before every launch
    float* attn_part_o;        // [n_kv * splits * GROUP * 128]
    float* attn_part_ml;       // [n_kv * splits * GROUP * 2]
    int* attn_tickets;         // [n_kv]
};

struct PhaseInfo {
    int type, layer, mode, map, K, n_kb, upr, n_rb;
    long long U, u0, u1;
};

__device__ __forceinline__ PhaseInfo phase_info(const MegaParams& P, int p) {
    PhaseInfo f;
    const int L = P.n_layers;
    if (p < 5 * L) { f.layer = p / 5; f.type = p % 5; } else { f.layer = L; f.type = PH_LMHEAD; }
    f.mode = MODE_SEQ;
    f.map = 0; f.K = P.H; f.n_rb = 0;
    switch (f.type) {
        case PH_QKV: f.map = f.layer * 4 + 0; f.n_rb = P.n_q + 2 * P.n_kv; break;
        case PH_OPROJ: f.map = f.layer * 4 + 1; f.K = P.n_q * MG_D; f.n_rb = (P.H + 127) / 128; break;
        case PH_GATEUP: f.map = f.layer * 4 + 2; f.mode = MODE_DUAL_SAMEX; f.n_rb = 2 * (P.I / 128); break;
        case PH_DOWN: f.map = f.layer * 4 + 3; f.mode = MODE_DUAL; f.K = P.I; f.n_rb = (P.H + 127) / 128; break;
        case PH_LMHEAD: f.map = L * 4; f.n_rb = (P.V + 127) / 128; break;
        default: break;
    }
    f.n_kb = f.K >> 6;
    f.upr = (f.mode == MODE_SEQ) ? (f.n_kb >> 1) : f.n_kb;   // units per row block (SEQ: two k-blocks per unit)
    f.U = (f.type == PH_ATTN) ? 0 : static_cast<long long>(f.n_rb) * f.upr;
    f.u0 = f.U * blockIdx.x / gridDim.x;
    f.u1 = f.U * (blockIdx.x + 1) / gridDim.x;
    return f;
}

__device__ __forceinline__ void mg_epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ float mg_epi_sum(float v, float* scratch) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[(threadIdx.x >> 5) - 4] = v;
    mg_epi_barrier();
    const float t = scratch[0] + scratch[1] + scratch[2] + scratch[3];
    mg_epi_barrier();
    return t;
}
// residual-stream reads must bypass L1: the element was last written by another CTA in an earlier phase
__device__ __forceinline__ float mg_ld_h(const __nv_bfloat16* p) {
    const unsigned short u = __ldcg(reinterpret_cast<const unsigned short*>(p));
    return __uint_as_float(static_cast<uint32_t>(u) << 16);
}
__device__ __forceinline__ void mg_tmem_ld1(uint32_t taddr, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

// Grid-wide barrier for the 128 epilogue threads of every CTA (all CTAs are co-resident: grid <= #SMs, 1 CTA/SM).
__device__ __forceinline__ void mg_grid_sync(unsigned int* counter, unsigned int target) {
    mg_epi_barrier();
    if (threadIdx.x == 128) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned int v;
        long long t0 = clock64();
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v < target && clock64() - t0 > VITA_MBAR_TIMEOUT_CYCLES) {
                printf("[vita] grid barrier timeout: block %d target %u seen %u\n", blockIdx.x, target, v);
                __trap();
            }
        } while (v < target);
    }
    mg_epi_barrier();
}

// xs = bf16(rmsnorm(h) * w) for K <= 4096 (128 threads x 4 chunks of 8), loads issued up front
__device__ __forceinline__ void mg_rmsnorm_to_xs(const __nv_bfloat16* h, const __nv_bfloat16* w, __nv_bfloat16* xs,
                                                 int K, float eps, float* scratch) {
    const int t = threadIdx.x - 128;
    uint4 hv[4], gv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = (t + j * 128) * 8;
        if (i < K) {
            hv[j] = __ldcg(reinterpret_cast<const uint4*>(h + i));
            gv[j] = __ldg(reinterpret_cast<const uint4*>(w + i));
        }
    }
    float ss = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if ((t + j * 128) * 8 < K) {
            const uint32_t a[4] = {hv[j].x, hv[j].y, hv[j].z, hv[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) ss += bf16_lo(a[e]) * bf16_lo(a[e]) + bf16_hi(a[e]) * bf16_hi(a[e]);
        }
    }
    const float inv = rsqrtf(mg_epi_sum(ss, scratch) / static_cast<float>(K) + eps);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = (t + j * 128) * 8;
        if (i < K) {
            const uint32_t a[4] = {hv[j].x, hv[j].y, hv[j].z, hv[j].w}, gg[4] = {gv[j].x, gv[j].y, gv[j].z, gv[j].w};
            uint4 o;
            uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                op[e] = pack_bf16(bf16_lo(a[e]) * inv * bf16_lo(gg[e]), bf16_hi(a[e]) * inv * bf16_hi(gg[e]));
            *reinterpret_cast<uint4*>(xs + i) = o;
        }
    }
}

struct MgSmem {
    __nv_bfloat16* xs;   // [4096]
    float* scratch;      // 64
    float* prep;         // 256: cos/sin row (QKV), router reduction (GATEUP)
    float* pair;         // 128
    float* attn_o;       // [4 warps][4 heads][128]
    float* attn_ml;      // [4 warps][4 heads][2]
    int* misc;           // [0..1] expert ids, [2] kv slot, [7] ticket flag
    float* tw;           // [2] routing weights of the layer
};

// ------------------------------------------------------------------------------------------------ attention phase
__device__ void mg_attention(const MegaParams& P, const MegaLayer& lw, const MgSmem& sm) {
    const int tid = threadIdx.x - 128;
    const int n_tasks = P.n_kv * P.splits;
    if (static_cast<int>(blockIdx.x) >= n_tasks) return;
    const int kvh = blockIdx.x / P.splits, split = blockIdx.x % P.splits;
    const int ctx = P.cur_pos[0] + 1;
    const int per = (ctx + P.splits - 1) / P.splits;
    const int k_begin = split * per;
    const int k_end = min(ctx, k_begin + per);
    const int lg = tid >> 3, sl = tid & 7, d0 = sl * 16;

    float q[MG_GROUP][16];
#pragma unroll
    for (int hh = 0; hh < MG_GROUP; ++hh) {
        const uint4* qp = reinterpret_cast<const uint4*>(P.q + (kvh * MG_GROUP + hh) * MG_D + d0);
        const uint4 a = __ldcg(qp), c = __ldcg(qp + 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { q[hh][2 * i] = bf16_lo(w[i]); q[hh][2 * i + 1] = bf16_hi(w[i]); }
    }
    float m[MG_GROUP], l[MG_GROUP], acc[MG_GROUP][16];
#pragma unroll
    for (int hh = 0; hh < MG_GROUP; ++hh) {
        m[hh] = -INFINITY; l[hh] = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[hh][i] = 0.0f;
    }
    uint4 nk0, nk1, nv0, nv1;
    auto issue = [&](int key) {
        if (key < k_end) {
            const int page = P.block_table[key / P.page_size];
            const long long slot = static_cast<long long>(page) * P.page_size + key % P.page_size;
            const uint4* kp = reinterpret_cast<const uint4*>(lw.k_cache + (slot * P.n_kv + kvh) * MG_D + d0);
            const uint4* vp = reinterpret_cast<const uint4*>(lw.v_cache + (slot * P.n_kv + kvh) * MG_D + d0);
            nk0 = __ldcg(kp); nk1 = __ldcg(kp + 1); nv0 = __ldcg(vp); nv1 = __ldcg(vp + 1);
        } else {
            nk0 = nk1 = nv0 = nv1 = make_uint4(0, 0, 0, 0);
        }
    };
    issue(k_begin + lg);
    for (int key0 = k_begin; key0 < k_end; key0 += 16) {
        const int key = key0 + lg;
        const bool valid = key < k_end;
        const uint4 ka = nk0, kc = nk1, va = nv0, vc = nv1;
        issue(key + 16);
        const uint32_t kw[8] = {ka.x, ka.y, ka.z, ka.w, kc.x, kc.y, kc.z, kc.w};
        const uint32_t vw[8] = {va.x, va.y, va.z, va.w, vc.x, vc.y, vc.z, vc.w};
        float kf[16], vf[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            kf[2 * i] = bf16_lo(kw[i]); kf[2 * i + 1] = bf16_hi(kw[i]);
            vf[2 * i] = bf16_lo(vw[i]); vf[2 * i + 1] = bf16_hi(vw[i]);
        }
#pragma unroll
        for (int hh = 0; hh < MG_GROUP; ++hh) {
            float dot = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i) dot += q[hh][i] * kf[i];
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            if (valid) {
                const float sc = dot * P.attn_scale_log2;
                const float m_new = fmaxf(m[hh], sc);
                const float alpha = exp2f(m[hh] - m_new);
                const float pr = exp2f(sc - m_new);
                m[hh] = m_new;
                l[hh] = l[hh] * alpha + pr;
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[hh][i] = acc[hh][i] * alpha + pr * vf[i];
            }
        }
    }
    // merge the 4 lane groups of each warp with shuffles (lane bits 3 and 4)
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
#pragma unroll
        for (int hh = 0; hh < MG_GROUP; ++hh) {
            const float mo = __shfl_xor_sync(0xffffffffu, m[hh], off);
            const float lo = __shfl_xor_sync(0xffffffffu, l[hh], off);
            const float mm = fmaxf(m[hh], mo);
            const float a = (m[hh] == -INFINITY) ? 0.0f : exp2f(m[hh] - mm);
            const float bw = (mo == -INFINITY) ? 0.0f : exp2f(mo - mm);
            l[hh] = l[hh] * a + lo * bw;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float ao = __shfl_xor_sync(0xffffffffu, acc[hh][i], off);
                acc[hh][i] = acc[hh][i] * a + ao * bw;
            }
            m[hh] = mm;
        }
    }
    const int w4 = tid >> 5;
    if ((tid & 31) < 8) {
#pragma unroll
        for (int hh = 0; hh < MG_GROUP; ++hh) {
            if (sl == 0) { sm.attn_ml[(w4 * MG_GROUP + hh) * 2] = m[hh]; sm.attn_ml[(w4 * MG_GROUP + hh) * 2 + 1] = l[hh]; }
#pragma unroll
            for (int i = 0; i < 16; ++i) sm.attn_o[(w4 * MG_GROUP + hh) * MG_D + d0 + i] = acc[hh][i];
        }
    }
    mg_epi_barrier();
    const int d = tid;
    const long long pbase = (static_cast<long long>(kvh) * P.splits + split) * MG_GROUP;
#pragma unroll
    for (int hh = 0; hh < MG_GROUP; ++hh) {
        float mm = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) mm = fmaxf(mm, sm.attn_ml[(w * MG_GROUP + hh) * 2]);
        float ll = 0.0f, oo = 0.0f;
        if (mm > -INFINITY) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float mw = sm.attn_ml[(w * MG_GROUP + hh) * 2];
                const float wt = (mw == -INFINITY) ? 0.0f : exp2f(mw - mm);
                ll += wt * sm.attn_ml[(w * MG_GROUP + hh) * 2 + 1];
                oo += wt * sm.attn_o[(w * MG_GROUP + hh) * MG_D + d];
            }
        }
        __stcg(&P.attn_part_o[(pbase + hh) * MG_D + d], oo);
        if (d == 0) { __stcg(&P.attn_part_ml[(pbase + hh) * 2], mm); __stcg(&P.attn_part_ml[(pbase + hh) * 2 + 1], ll); }
    }
    mg_epi_barrier();
    if (tid == 0) {
        int t;
        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(t) : "l"(P.attn_tickets + kvh) : "memory");
        sm.misc[7] = (t == P.splits - 1);
        if (t == P.splits - 1) P.attn_tickets[kvh] = 0;
    }
    mg_epi_barrier();
    const bool last = sm.misc[7] != 0;
    mg_epi_barrier();
    if (!last) return;
    const long long sbase = static_cast<long long>(kvh) * P.splits * MG_GROUP;
#pragma unroll
    for (int hh = 0; hh < MG_GROUP; ++hh) {
        float ms[MG_MAXS], ls[MG_MAXS], os[MG_MAXS];
#pragma unroll
        for (int s = 0; s < MG_MAXS; ++s) {
            if (s < P.splits) {
                ms[s] = __ldcg(&P.attn_part_ml[(sbase + s * MG_GROUP + hh) * 2]);
                ls[s] = __ldcg(&P.attn_part_ml[(sbase + s * MG_GROUP + hh) * 2 + 1]);
                os[s] = __ldcg(&P.attn_part_o[(sbase + s * MG_GROUP + hh) * MG_D + d]);
            } else {
                ms[s] = -INFINITY; ls[s] = 0.0f; os[s] = 0.0f;
            }
        }
        float mm = -INFINITY;
#pragma unroll
        for (int s = 0; s < MG_MAXS; ++s) mm = fmaxf(mm, ms[s]);
        float ll = 0.0f, oo = 0.0f;
#pragma unroll
        for (int s = 0; s < MG_MAXS; ++s) {
            const float wt = (ms[s] == -INFINITY) ? 0.0f : exp2f(ms[s] - mm);
            ll += wt * ls[s];
            oo += wt * os[s];
        }
        P.attn[(kvh * MG_GROUP + hh) * MG_D + d] = __float2bfloat16(ll > 0.0f ? oo / ll : 0.0f);
    }
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(MG_THREADS, 1)
decode_mega_kernel(const MegaParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sX = sA + MG_STAGES * MG_STAGE_A;
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(sX + MG_STAGES * MG_STAGE_X);
    uint8_t* tail = reinterpret_cast<uint8_t*>(xs) + 4096 * 2;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + MG_STAGES;
    uint64_t* acc_full = empty_bar + MG_STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* x_ready = acc_empty + 2;
    uint64_t* ids_ready = x_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ids_ready + 1);
    float* fbase = reinterpret_cast<float*>(tmem_slot + 2);
    MgSmem sm;
    sm.xs = xs;
    sm.scratch = fbase;
    sm.prep = fbase + 64;
    sm.pair = sm.prep + 256;
    sm.attn_o = sm.pair + 128;
    sm.attn_ml = sm.attn_o + 4 * MG_GROUP * MG_D;
    sm.tw = sm.attn_ml + 4 * MG_GROUP * 2;
    sm.misc = reinterpret_cast<int*>(sm.tw + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_phases = 5 * P.n_layers + 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < MG_STAGES; ++i) {
            mbar_init(&full_bar[i], 2);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);
        }
        mbar_init(x_ready, 1);
        mbar_init(ids_ready, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 64);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < MG_STAGES * MG_STAGE_X / 16; i += MG_THREADS) reinterpret_cast<uint4*>(sX)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer: never waits for activations
            int stage = 0;
            uint32_t phase = 0, ids_par = 0;
            for (int p = 0; p < n_phases; ++p) {
                const PhaseInfo f = phase_info(P, p);
                if (f.type == PH_ATTN) continue;
                if (f.type == PH_GATEUP) { mbar_wait(ids_ready, ids_par, 31); ids_par ^= 1; }
                const CUtensorMap* tm = P.maps + f.map;
                const int e0 = sm.misc[0], e1 = sm.misc[1];
                const int nb = P.I / 128;
                for (long long u = f.u0; u < f.u1; ++u) {
                    const int rb = static_cast<int>(u / f.upr), j = static_cast<int>(u % f.upr);
                    int row0, row1, k0, k1;
                    if (f.mode == MODE_SEQ) {
                        row0 = row1 = rb * 128; k0 = 2 * j * 64; k1 = k0 + 64;
                    } else if (f.mode == MODE_DUAL_SAMEX) {
                        const int k = rb / nb, jb = rb % nb;
                        row0 = (k == 0 ? e0 : e1) * 2 * P.I + jb * 128; row1 = row0 + P.I; k0 = k1 = j * 64;
                    } else {
                        row0 = e0 * P.H + rb * 128; row1 = e1 * P.H + rb * 128; k0 = k1 = j * 64;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1, 32);
                    mbar_arrive_expect_tx(&full_bar[stage], MG_STAGE_A);
                    tma_load_2d(sA + stage * MG_STAGE_A, tm, &full_bar[stage], k0, row0);
                    tma_load_2d(sA + stage * MG_STAGE_A + MG_A_BYTES, tm, &full_bar[stage], k1, row1);
                    if (++stage == MG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = umma_idesc_bf16(128, 16);
            const uint32_t sA_addr = smem_u32(sA), sX_addr = smem_u32(sX);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int p = 0; p < n_phases; ++p) {
                const PhaseInfo f = phase_info(P, p);
                if (f.type == PH_ATTN) continue;
                long long u = f.u0;
                while (u < f.u1) {
                    const int rb = static_cast<int>(u / f.upr);
                    long long seg_end = static_cast<long long>(rb + 1) * f.upr;
                    if (seg_end > f.u1) seg_end = f.u1;
                    mbar_wait(&acc_empty[acc], acc_phase ^ 1, 33);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 32;
                    for (long long v = u; v < seg_end; ++v) {
                        mbar_wait(&full_bar[stage], phase, 34);
                        tc_fence_after();
#pragma unroll
                        for (int part = 0; part < 2; ++part) {
                            const uint64_t da = umma_desc_k_sw128(sA_addr + stage * MG_STAGE_A + part * MG_A_BYTES);
                            const int xi = (f.mode == MODE_DUAL_SAMEX) ? 0 : part;
                            const uint64_t dx = umma_desc_k_sw128(sX_addr + stage * MG_STAGE_X + xi * MG_X_BYTES);
                            const uint32_t dcol = d_tmem + ((f.mode == MODE_SEQ) ? 0 : part * 16);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const bool first = (v == u) && k == 0 && (f.mode != MODE_SEQ || part == 0);
                                tc_mma_bf16(dcol, da + 2 * k, dx + 2 * k, idesc, first ? 0u : 1u);
                            }
                        }
                        tc_commit(&empty_bar[stage]);
                        if (++stage == MG_STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&acc_full[acc]);
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                    u = seg_end;
                }
            }
        }
    } else if (warp == 3) {
        // ---------------------------------------------------------------- x-tile writer
        int stage = 0;
        uint32_t phase = 0, xr_par = 0;
        const int px = lane >> 3, c = lane & 7;
        for (int p = 0; p < n_phases; ++p) {
            const PhaseInfo f = phase_info(P, p);
            if (f.type == PH_ATTN) continue;
            mbar_wait(x_ready, xr_par, 35);
            xr_par ^= 1;
            if (f.mode == MODE_DUAL) {
                constexpr int LOOK = 4;
                const bool active = lane < 16;
                uint4 nxt[LOOK];
#pragma unroll
                for (int j = 0; j < LOOK; ++j)
                    if (active && f.u0 + j < f.u1)
                        nxt[j] = __ldcg(reinterpret_cast<const uint4*>(P.act + static_cast<long long>(px) * P.I +
                                                                       ((f.u0 + j) % f.upr) * 64 + c * 8));
                for (long long u = f.u0; u < f.u1; u += LOOK) {
#pragma unroll
                    for (int j = 0; j < LOOK; ++j) {
                        if (u + j < f.u1) {
                            mbar_wait(&empty_bar[stage], phase ^ 1, 36);
                            if (active) {
                                *reinterpret_cast<uint4*>(sX + stage * MG_STAGE_X + px * MG_X_BYTES + c * 16) = nxt[j];
                                if (u + j + LOOK < f.u1)
                                    nxt[j] = __ldcg(reinterpret_cast<const uint4*>(
                                        P.act + static_cast<long long>(px) * P.I + ((u + j + LOOK) % f.upr) * 64 + c * 8));
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&full_bar[stage]);
                            if (++stage == MG_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            } else {
                for (long long u = f.u0; u < f.u1; ++u) {
                    const int j = static_cast<int>(u % f.upr);
                    mbar_wait(&empty_bar[stage], phase ^ 1, 36);
                    if (f.mode == MODE_SEQ) {
                        if (lane < 16)
                            *reinterpret_cast<uint4*>(sX + stage * MG_STAGE_X + px * MG_X_BYTES + c * 16) =
                                *reinterpret_cast<const uint4*>(xs + (2 * j + px) * 64 + c * 8);
                    } else if (lane < 8) {
                        *reinterpret_cast<uint4*>(sX + stage * MG_STAGE_X + c * 16) =
                            *reinterpret_cast<const uint4*>(xs + j * 64 + c * 8);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[stage]);
                    if (++stage == MG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- prologues, epilogues, attention
        const int t = threadIdx.x - 128;
        const int quad = warp - 4, row = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        unsigned int bar_target = 0;
        float best_v = -INFINITY;
        int best_i = 0x7fffffff;
        for (int p = 0; p < n_phases; ++p) {
            const PhaseInfo f = phase_info(P, p);
            if (p > 0) { bar_target += gridDim.x; mg_grid_sync(P.grid_bar, bar_target); }
            const MegaLayer* lw = (f.layer < P.n_layers) ? &P.layers[f.layer] : nullptr;
            // ---- prologue
            if (f.type == PH_ATTN) {
                mg_attention(P, *lw, sm);
                continue;
            } else if (f.type == PH_QKV) {
                const int pos = P.cur_pos[0];
                sm.prep[t] = P.cos_sin[static_cast<long long>(pos) * 128 + t];
                if (t == 0) sm.misc[2] = P.block_table[pos / P.page_size] * P.page_size + pos % P.page_size;
                mg_rmsnorm_to_xs(P.h, lw->ln1, xs, P.H, P.eps, sm.scratch);
            } else if (f.type == PH_LMHEAD) {
                mg_rmsnorm_to_xs(P.h, P.final_norm, xs, P.H, P.eps, sm.scratch);
            } else if (f.type == PH_OPROJ) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = (t + j * 128) * 8;
                    if (i < f.K) *reinterpret_cast<uint4*>(xs + i) = __ldcg(reinterpret_cast<const uint4*>(P.attn + i));
                }
            } else if (f.type == PH_GATEUP) {
                // post-attention RMSNorm + router (top-2 of 8), every CTA recomputes it
                const int K = P.H;
                float part[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) part[e] = 0.0f;
                uint4 hv[4], gv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = (t + j * 128) * 8;
                    if (i < K) {
                        hv[j] = __ldcg(reinterpret_cast<const uint4*>(P.h + i));
                        gv[j] = __ldg(reinterpret_cast<const uint4*>(lw->ln2 + i));
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = (t + j * 128) * 8;
                    if (i < K) {
                        uint4 ge[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) ge[e] = __ldg(reinterpret_cast<const uint4*>(lw->gate + static_cast<long long>(e) * K + i));
                        const uint32_t a[4] = {hv[j].x, hv[j].y, hv[j].z, hv[j].w}, gg[4] = {gv[j].x, gv[j].y, gv[j].z, gv[j].w};
                        float xw[8];
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const float lo = bf16_lo(a[qd]), hi = bf16_hi(a[qd]);
                            part[0] += lo * lo + hi * hi;
                            xw[2 * qd] = lo * bf16_lo(gg[qd]);
                            xw[2 * qd + 1] = hi * bf16_hi(gg[qd]);
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const uint32_t w[4] = {ge[e].x, ge[e].y, ge[e].z, ge[e].w};
#pragma unroll
                            for (int qd = 0; qd < 4; ++qd)
                                part[1 + e] += xw[2 * qd] * bf16_lo(w[qd]) + xw[2 * qd + 1] * bf16_hi(w[qd]);
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 9; ++e) part[e] = warp_sum(part[e]);
                if (lane == 0)
#pragma unroll
                    for (int e = 0; e < 9; ++e) sm.prep[quad * 9 + e] = part[e];
                mg_epi_barrier();
                float tot[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) tot[e] = sm.prep[e] + sm.prep[9 + e] + sm.prep[18 + e] + sm.prep[27 + e];
                const float inv = rsqrtf(tot[0] / static_cast<float>(K) + P.eps);
                float pr[8], mx = -INFINITY, sum = 0.0f;
#pragma unroll
                for (int e = 0; e < 8; ++e) { pr[e] = tot[1 + e] * inv; mx = fmaxf(mx, pr[e]); }
#pragma unroll
                for (int e = 0; e < 8; ++e) { pr[e] = expf(pr[e] - mx); sum += pr[e]; }
                int e0 = 0;
#pragma unroll
                for (int e = 1; e < 8; ++e) if (pr[e] > pr[e0]) e0 = e;
                int e1 = (e0 == 0) ? 1 : 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) if (e != e0 && pr[e] > pr[e1]) e1 = e;
                mg_epi_barrier();   // everyone has read prep before it is reused
                if (t == 0) {
                    const float p0 = pr[e0] / sum, p1 = pr[e1] / sum, den = p0 + p1;
                    sm.misc[0] = e0;
                    sm.misc[1] = e1;
                    sm.tw[0] = p0 / den;
                    sm.tw[1] = p1 / den;
                    mbar_arrive(ids_ready);   // the producer may now stream this layer's expert weights
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = (t + j * 128) * 8;
                    if (i < K) {
                        const uint32_t a[4] = {hv[j].x, hv[j].y, hv[j].z, hv[j].w}, gg[4] = {gv[j].x, gv[j].y, gv[j].z, gv[j].w};
                        uint4 o;
                        uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd)
                            op[qd] = pack_bf16(bf16_lo(a[qd]) * inv * bf16_lo(gg[qd]), bf16_hi(a[qd]) * inv * bf16_hi(gg[qd]));
                        *reinterpret_cast<uint4*>(xs + i) = o;
                    }
                }
            }
            mg_epi_barrier();
            if (t == 0) mbar_arrive(x_ready);

            // ---- segments
            long long u = f.u0;
            while (u < f.u1) {
                const int rb = static_cast<int>(u / f.upr);
                const long long rb_begin = static_cast<long long>(rb) * f.upr, rb_end = rb_begin + f.upr;
                const long long seg_end = rb_end < f.u1 ? rb_end : f.u1;
                mbar_wait(&acc_full[acc], acc_phase, 37);
                tc_fence_after();
                float v[2];
                uint32_t r0, r1;
                mg_tmem_ld1(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 32, r0);
                mg_tmem_ld1(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 32 + 16, r1);
                tmem_ld_wait();
                v[0] = __uint_as_float(r0);
                v[1] = __uint_as_float(r1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;

                bool do_finish = true;
                if (u != rb_begin || seg_end != rb_end) {
                    int c_first = blockIdx.x;
                    while (c_first > 0 && f.U * c_first / gridDim.x > rb_begin) --c_first;
                    int c_last = blockIdx.x;
                    while (c_last + 1 < static_cast<int>(gridDim.x) && f.U * (c_last + 1) / gridDim.x < rb_end) ++c_last;
                    const int slot = blockIdx.x - c_first, n_contrib = c_last - c_first + 1;
                    float* base = P.scratch + static_cast<long long>(rb) * MG_SLOTS * 2 * 128;
                    __stcg(base + (slot * 2 + 0) * 128 + row, v[0]);
                    __stcg(base + (slot * 2 + 1) * 128 + row, v[1]);
                    mg_epi_barrier();
                    if (t == 0) {
                        int tk;
                        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(tk) : "l"(P.tickets + rb) : "memory");
                        sm.misc[7] = (tk == n_contrib - 1);
                        if (tk == n_contrib - 1) P.tickets[rb] = 0;
                    }
                    mg_epi_barrier();
                    do_finish = sm.misc[7] != 0;
                    if (do_finish) {
                        float s0 = 0.0f, s1 = 0.0f;
                        for (int qd = 0; qd < n_contrib; ++qd) {
                            s0 += __ldcg(base + (qd * 2 + 0) * 128 + row);
                            s1 += __ldcg(base + (qd * 2 + 1) * 128 + row);
                        }
                        v[0] = s0;
                        v[1] = s1;
                    }
                    mg_epi_barrier();
                }
                if (do_finish) {
                    if (f.type == PH_QKV) {
                        const float x = __bfloat162float(__float2bfloat16(v[0]));
                        float o = x;
                        if (rb < P.n_q + P.n_kv) {
                            sm.pair[row] = x;
                            mg_epi_barrier();
                            const float partner = sm.pair[row ^ 64];
                            const int j = row & 63;
                            const float cs = sm.prep[j], sn = sm.prep[64 + j];
                            o = (row < 64) ? x * cs - partner * sn : x * cs + partner * sn;
                            mg_epi_barrier();
                        }
                        if (rb < P.n_q) {
                            P.q[rb * 128 + row] = __float2bfloat16(o);
                        } else {
                            const long long slot = sm.misc[2];
                            const bool is_k = rb < P.n_q + P.n_kv;
                            const int kvh = is_k ? rb - P.n_q : rb - P.n_q - P.n_kv;
                            ((is_k ? lw->k_cache : lw->v_cache) + (slot * P.n_kv + kvh) * 128)[row] = __float2bfloat16(o);
                        }
                    } else if (f.type == PH_OPROJ) {
                        const int r = rb * 128 + row;
                        if (r < P.H) P.h[r] = __float2bfloat16(mg_ld_h(P.h + r) + v[0]);
                    } else if (f.type == PH_GATEUP) {
                        const int nb = P.I / 128, k = rb / nb, jb = rb % nb;
                        P.act[static_cast<long long>(k) * P.I + jb * 128 + row] = __float2bfloat16(silu(v[0]) * v[1]);
                    } else if (f.type == PH_DOWN) {
                        const int r = rb * 128 + row;
                        if (r < P.H) P.h[r] = __float2bfloat16(mg_ld_h(P.h + r) + sm.tw[0] * v[0] + sm.tw[1] * v[1]);
                    } else {  // PH_LMHEAD
                        const int r = rb * 128 + row;
                        if (r < P.V) {
                            const __nv_bfloat16 lg = __float2bfloat16(v[0]);
                            if (P.logits) P.logits[r] = lg;
                            const float fv = __bfloat162float(lg);
                            if (fv > best_v || (fv == best_v && r < best_i)) { best_v = fv; best_i = r; }
                        }
                    }
                }
                u = seg_end;
            }
        }
        // arg-max of the bf16 logits: warp reduce, one 64-bit atomicMax per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
        }
        if (lane == 0 && best_i != 0x7fffffff) {
            uint32_t uu = __float_as_uint(best_v);
            uu = (uu & 0x80000000u) ? ~uu : (uu | 0x80000000u);
            atomicMax(P.best, (static_cast<unsigned long long>(uu) << 32) |
                                  static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(best_i)));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

constexpr int mega_smem_bytes() {
    return MG_STAGES * (MG_STAGE_A + MG_STAGE_X) + 4096 * 2 + (2 * MG_STAGES + 6) * 8 + 8 +
           (64 + 256 + 128 + 4 * MG_GROUP * MG_D + 4 * MG_GROUP * 2 + 2 + 8) * 4 + 1024 + 256;
}

}  // namespace vita

using namespace vita;

// Device-resident description of the model for the single-kernel decode step.
//   maps_out:   device buffer of (n_layers * 4 + 1) * sizeof(CUtensorMap) bytes (64-byte aligned)
//   layers_out: device buffer of n_layers * sizeof(MegaLayer) bytes
extern "C" int64_t vita_mega_maps_bytes(int64_t n_layers) { return (n_layers * 4 + 1) * static_cast<int64_t>(sizeof(CUtensorMap)); }
extern "C" int64_t vita_mega_layers_bytes(int64_t n_layers) { return n_layers * static_cast<int64_t>(sizeof(MegaLayer)); }
extern "C" int64_t vita_mega_workspace_floats(int64_t max_row_blocks) { return max_row_blocks * MG_SLOTS * 2 * 128; }

extern "C" int vita_mega_build(void* maps_out, void* layers_out, int64_t n_layers, const void* const* wqkv,
                               const void* const* wo, const void* const* w13, const void* const* w2,
                               const void* const* ln1, const void* const* ln2, const void* const* gate,
                               void* const* k_cache, void* const* v_cache, const void* lm_head, int64_t H, int64_t I,
                               int64_t E, int64_t n_q, int64_t n_kv, int64_t V) {
    VITA_REQUIRE(H % 128 == 0 && H <= 4096 && I % 128 == 0, "single-kernel decode needs H % 128 == 0, H <= 4096, I % 128 == 0");
    VITA_REQUIRE((n_q * 128) % 128 == 0 && n_q * 128 <= 4096, "n_q * 128 must be <= 4096");
    std::vector<CUtensorMap> maps(static_cast<size_t>(n_layers * 4 + 1));
    std::vector<MegaLayer> layers(static_cast<size_t>(n_layers));
    const uint32_t box[2] = {64, 128};
    auto mk = [&](CUtensorMap* m, const void* base, uint64_t rows, uint64_t K) {
        const uint64_t dims[2] = {K, rows};
        const uint64_t strides[1] = {K * 2};
        return make_tensor_map_bf16(m, base, 2, dims, strides, box, true);
    };
    for (int64_t l = 0; l < n_layers; ++l) {
        int rc = mk(&maps[l * 4 + 0], wqkv[l], static_cast<uint64_t>((n_q + 2 * n_kv) * 128), static_cast<uint64_t>(H));
        if (!rc) rc = mk(&maps[l * 4 + 1], wo[l], static_cast<uint64_t>(H), static_cast<uint64_t>(n_q * 128));
        if (!rc) rc = mk(&maps[l * 4 + 2], w13[l], static_cast<uint64_t>(E * 2 * I), static_cast<uint64_t>(H));
        if (!rc) rc = mk(&maps[l * 4 + 3], w2[l], static_cast<uint64_t>(E * H), static_cast<uint64_t>(I));
        if (rc) return rc;
        layers[l] = MegaLayer{BF16C(ln1[l]), BF16C(ln2[l]), BF16C(gate[l]), static_cast<__nv_bfloat16*>(k_cache[l]),
                              static_cast<__nv_bfloat16*>(v_cache[l])};
    }
    int rc = mk(&maps[n_layers * 4], lm_head, static_cast<uint64_t>(V), static_cast<uint64_t>(H));
    if (rc) return rc;
    rc = check_cuda(cudaMemcpy(maps_out, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice), "mega maps");
    if (rc) return rc;
    return check_cuda(cudaMemcpy(layers_out, layers.data(), layers.size() * sizeof(MegaLayer), cudaMemcpyHostToDevice),
                      "mega layers");
}

extern "C" int vita_mega_decode_step(const void* maps, const void* layers, int64_t n_layers, const void* final_norm,
                                     void* h, void* q, void* attn, void* act, void* logits, uint64_t* best,
                                     const float* cos_sin, const int32_t* cur_pos, const int32_t* block_table,
                                     float* scratch, int32_t* tickets, uint32_t* grid_bar, float* attn_part_o,
                                     float* attn_part_ml, int32_t* attn_tickets, int64_t H, int64_t I, int64_t n_q,
                                     int64_t n_kv, int64_t V, int64_t page_size, int64_t max_pages, int64_t splits,
                                     float eps, float attn_scale, void* stream) {
    VITA_REQUIRE(splits >= 1 && splits <= MG_MAXS && n_kv * splits <= num_sms(), "attention splits");
    VITA_REQUIRE(n_q == n_kv * MG_GROUP, "GQA group must be 4");
    MegaParams P{};
    P.maps = static_cast<const CUtensorMap*>(maps);
    P.layers = static_cast<const MegaLayer*>(layers);
    P.n_layers = (int)n_layers;
    P.final_norm = BF16C(final_norm);
    P.H = (int)H; P.I = (int)I; P.n_q = (int)n_q; P.n_kv = (int)n_kv; P.V = (int)V;
    P.page_size = (int)page_size; P.max_pages = (int)max_pages; P.splits = (int)splits;
    P.eps = eps; P.attn_scale_log2 = attn_scale * 1.4426950408889634f;
    P.h = static_cast<__nv_bfloat16*>(h); P.q = static_cast<__nv_bfloat16*>(q);
    P.attn = static_cast<__nv_bfloat16*>(attn); P.act = static_cast<__nv_bfloat16*>(act);
    P.logits = static_cast<__nv_bfloat16*>(logits);
    P.best = reinterpret_cast<unsigned long long*>(best);
    P.cos_sin = cos_sin; P.cur_pos = cur_pos; P.block_table = block_table;
    P.scratch = scratch; P.tickets = tickets; P.grid_bar = grid_bar;
    P.attn_part_o = attn_part_o; P.attn_part_ml = attn_part_ml; P.attn_tickets = attn_tickets;
    constexpr int smem = mega_smem_bytes();
    static bool configured = false;
    if (!configured) {
        int rc = check_cuda(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                            "mega smem");
        if (rc) return rc;
        configured = true;
    }
    // every CTA must get at least ceil(n_kb / 6) units in every phase (bounded contributors per row block): the
    // smallest phase is the o-projection / qkv; with the full model geometry all 148 SMs qualify
    int grid = num_sms();
    const long long u_min = static_cast<long long>((H + 127) / 128) * ((n_q * 128) / 128);   // OPROJ units (SEQ)
    const int min_units = static_cast<int>(((n_q * 128 / 64) / 2 + 5) / 6);
    if (u_min / (min_units > 0 ? min_units : 1) < grid) grid = static_cast<int>(u_min / (min_units > 0 ? min_units : 1));
    if (grid < static_cast<int>(n_kv * splits)) {
        set_last_error("vita_mega_decode_step: model too small for the single-kernel step (use the per-kernel path)");
        return VITA_ERR_INVALID;
    }
    decode_mega_kernel<<<grid, MG_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(P);
    return check_launch("decode_mega_kernel");
}
