// Attention kernels.
//
// (1) flash_fwd_kernel<DQK, DV>: FlashAttention-style fused softmax(Q K^T * scale) V for the three prefill-shaped
//     attentions of the path: Mixtral causal GQA (128/128), InternViT non-causal (64/64), Whale rel-pos with the two
//     score terms folded into one contraction over [k | p] (128/64).  K/V tiles are staged through XOR-swizzled shared
//     memory with cp.async double buffering; the contractions run on mma.sync m16n8k16 bf16 (legacy tensor path --
//     attention is < 1% of the prefill FLOPs at the BASELINE sequence lengths; the tcgen05 port is DESIGN.md "next").
// (2) decode_attn_kernel: single-query paged-KV attention for greedy decode.  One CTA per (kv head, context split);
//     the 4 query heads of a GQA group share every K/V load; 8-lane groups own one key each and reduce with warp
//     shuffles.  Under programmatic dependent launch the cached K/V rows are fetched ahead of the dependency wait
//     (only q and the newest row come from the QKV kernel).  Splits are merged all-to-all through tagged 64-bit
//     words, each CTA owning a slice of the outputs (TAGGED = true; latency chain of one publish + one poll), or by
//     the last CTA to take a ticket (TAGGED = false: odd split counts, grids too large to be resident at once).
//
// Reference call sites: flash_attn_varlen_qkvpacked_func (internvit/flash_attention.py:61) / naive softmax
// (modeling_intern_vit.py:170-174); whale attention.py:391-415; transformers sdpa/eager attention
// (modeling_mixtral.py:269-292); vLLM paged Attention (web_demo/vllm_tools/vllm_file/mixtral.py:484-501).
#include "common.h"
#include "ptx.cuh"

namespace vita {

struct AttnParams {
    const __nv_bfloat16* q;
    const __nv_bfloat16* k;
    const __nv_bfloat16* v;
    __nv_bfloat16* o;
    long long q_bs, q_ts, q_hs;  // batch / token / head strides in elements
    long long k_bs, k_ts, k_hs;
    long long v_bs, v_ts, v_hs;
    long long o_bs, o_ts, o_hs;
    int group;         // query heads per kv head
    int Sq, Skv;
    const int* kv_lens;  // [B] valid keys per batch entry, or nullptr
    int causal;
    float scale_log2;  // softmax scale * log2(e)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `chunk` of row `row` in a [rows][D] bf16 tile with XOR swizzle
template <int D>
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
    return static_cast<uint32_t>(row * D * 2 + ((chunk ^ (row & 7)) << 4));
}

template <int D, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile_async(uint32_t smem_base, const __nv_bfloat16* g, long long tok_stride,
                                                int row0, int n_valid_rows) {
    constexpr int CPR = D / 8;  // chunks per row
    constexpr int TOTAL = ROWS * CPR;
#pragma unroll
    for (int i = 0; i < TOTAL / THREADS; ++i) {
        const int c = threadIdx.x + i * THREADS;
        const int r = c / CPR, ch = c % CPR;
        const bool ok = (row0 + r) < n_valid_rows;
        const __nv_bfloat16* src = g + static_cast<long long>(ok ? (row0 + r) : 0) * tok_stride + ch * 8;
        cp_async16(smem_base + swz<D>(r, ch), src, ok);
    }
}

// BM query rows per CTA (BM / 16 warps); K/V tiles of 64 keys.  BM = 128 halves the K/V shared-memory traffic per FLOP
// (used for long sequences), BM = 64 keeps enough CTAs in flight for short ones.
template <int DQK, int DV, int BM>
__global__ void __launch_bounds__(BM * 2)
flash_fwd_kernel(const AttnParams p) {
    constexpr int THREADS = BM * 2;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = sQ + BM * DQK * 2;
    const uint32_t sV = sK + 2 * 64 * DQK * 2;

    const int m_blk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int kvh = head / p.group;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int kv_len = p.kv_lens ? p.kv_lens[b] : p.Skv;
    if (kv_len > p.Skv) kv_len = p.Skv;
    const int q0 = m_blk * BM;
    int n_tiles = (kv_len + 63) / 64;
    if (p.causal && n_tiles > (q0 + BM + 63) / 64) n_tiles = (q0 + BM + 63) / 64;

    const __nv_bfloat16* qg = p.q + b * p.q_bs + head * p.q_hs;
    const __nv_bfloat16* kg = p.k + b * p.k_bs + kvh * p.k_hs;
    const __nv_bfloat16* vg = p.v + b * p.v_bs + kvh * p.v_hs;

    load_tile_async<DQK, BM, THREADS>(sQ, qg, p.q_ts, q0, p.Sq);
    if (n_tiles > 0) {
        load_tile_async<DQK, 64, THREADS>(sK, kg, p.k_ts, 0, kv_len);
        load_tile_async<DV, 64, THREADS>(sV, vg, p.v_ts, 0, kv_len);
    }
    cp_async_commit();

    float o_acc[DV / 8][4];
#pragma unroll
    for (int i = 0; i < DV / 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) o_acc[i][e] = 0.0f;
    float m_i[2] = {-INFINITY, -INFINITY};
    float l_i[2] = {0.0f, 0.0f};
    uint32_t q_frag[DQK / 16][4];

    const int g = lane >> 2, tq = lane & 3;
    const int ld_i = lane >> 3, ld_r = lane & 7;

    for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        if (j + 1 < n_tiles) {
            load_tile_async<DQK, 64, THREADS>(sK + (buf ^ 1) * 64 * DQK * 2, kg, p.k_ts, (j + 1) * 64, kv_len);
            load_tile_async<DV, 64, THREADS>(sV + (buf ^ 1) * 64 * DV * 2, vg, p.v_ts, (j + 1) * 64, kv_len);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (j == 0) {
#pragma unroll
            for (int ks = 0; ks < DQK / 16; ++ks)
                ldsm_x4(sQ + swz<DQK>(warp * 16 + (ld_i & 1) * 8 + ld_r, ks * 2 + (ld_i >> 1)), q_frag[ks][0],
                        q_frag[ks][1], q_frag[ks][2], q_frag[ks][3]);
        }
        const uint32_t kb = sK + buf * 64 * DQK * 2;
        const uint32_t vb = sV + buf * 64 * DV * 2;

        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[i][e] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < DQK / 16; ++ks) {
#pragma unroll
            for (int nt2 = 0; nt2 < 4; ++nt2) {
                uint32_t r0, r1, r2, r3;
                ldsm_x4(kb + swz<DQK>(nt2 * 16 + (ld_i >> 1) * 8 + ld_r, ks * 2 + (ld_i & 1)), r0, r1, r2, r3);
                mma16816(s[nt2 * 2], q_frag[ks], r0, r1);
                mma16816(s[nt2 * 2 + 1], q_frag[ks], r2, r3);
            }
        }

        // scale, mask, online softmax (base 2)
        const int row_lo = q0 + warp * 16 + g;
        float rmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = j * 64 + nt * 8 + tq * 2 + (e & 1);
                const int row = row_lo + (e >> 1) * 8;
                float v = s[nt][e] * p.scale_log2;
                if (col >= kv_len || (p.causal && col > row)) v = -INFINITY;
                s[nt][e] = v;
                rmax[e >> 1] = fmaxf(rmax[e >> 1], v);
            }
        }
        float alpha[2], m_safe[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 1));
            rmax[r] = fmaxf(rmax[r], __shfl_xor_sync(0xffffffffu, rmax[r], 2));
            const float m_new = fmaxf(m_i[r], rmax[r]);
            m_safe[r] = (m_new == -INFINITY) ? 0.0f : m_new;
            alpha[r] = exp2f(m_i[r] - m_safe[r]);
            m_i[r] = m_new;
        }
        float rsum[2] = {0.0f, 0.0f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float pv = exp2f(s[nt][e] - m_safe[e >> 1]);
                s[nt][e] = pv;
                rsum[e >> 1] += pv;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_i[r] = l_i[r] * alpha[r] + rsum[r];
#pragma unroll
        for (int dt = 0; dt < DV / 8; ++dt) {
            o_acc[dt][0] *= alpha[0];
            o_acc[dt][1] *= alpha[0];
            o_acc[dt][2] *= alpha[1];
            o_acc[dt][3] *= alpha[1];
        }
        // O += P V
#pragma unroll
        for (int ks2 = 0; ks2 < 4; ++ks2) {
            uint32_t a[4];
            a[0] = pack_bf16(s[ks2 * 2][0], s[ks2 * 2][1]);
            a[1] = pack_bf16(s[ks2 * 2][2], s[ks2 * 2][3]);
            a[2] = pack_bf16(s[ks2 * 2 + 1][0], s[ks2 * 2 + 1][1]);
            a[3] = pack_bf16(s[ks2 * 2 + 1][2], s[ks2 * 2 + 1][3]);
#pragma unroll
            for (int dt2 = 0; dt2 < DV / 16; ++dt2) {
                uint32_t r0, r1, r2, r3;
                ldsm_x4_t(vb + swz<DV>(ks2 * 16 + (ld_i & 1) * 8 + ld_r, dt2 * 2 + (ld_i >> 1)), r0, r1, r2, r3);
                mma16816(o_acc[dt2 * 2], a, r0, r1);
                mma16816(o_acc[dt2 * 2 + 1], a, r2, r3);
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
    }
    __nv_bfloat16* og = p.o + b * p.o_bs + head * p.o_hs;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + g + r * 8;
        if (row >= p.Sq) continue;
        const float inv = l_i[r] > 0.0f ? 1.0f / l_i[r] : 0.0f;
        __nv_bfloat16* orow = og + static_cast<long long>(row) * p.o_ts;
#pragma unroll
        for (int dt = 0; dt < DV / 8; ++dt) {
            const uint32_t packed = pack_bf16(o_acc[dt][r * 2] * inv, o_acc[dt][r * 2 + 1] * inv);
            *reinterpret_cast<uint32_t*>(orow + dt * 8 + tq * 2) = packed;
        }
    }
}

template <int DQK, int DV, int BM>
static int launch_flash_bm(const AttnParams& p, int B, int Hq, cudaStream_t st) {
    constexpr int smem_bytes = BM * DQK * 2 + 2 * 64 * DQK * 2 + 2 * 64 * DV * 2;
    static bool configured = false;
    auto kern = flash_fwd_kernel<DQK, DV, BM>;
    if (!configured) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes),
                            "cudaFuncSetAttribute(flash smem)");
        if (rc) return rc;
        configured = true;
    }
    dim3 grid((p.Sq + BM - 1) / BM, Hq, B);
    kern<<<grid, BM * 2, smem_bytes, st>>>(p);
    return check_launch("flash_fwd_kernel");
}

template <int DQK, int DV>
static int launch_flash(const AttnParams& p, int B, int Hq, cudaStream_t st) {
    // 128-row CTAs once they still fill the machine twice over
    const long long ctas128 = static_cast<long long>((p.Sq + 127) / 128) * Hq * B;
    if (ctas128 >= 2ll * num_sms()) return launch_flash_bm<DQK, DV, 128>(p, B, Hq, st);
    return launch_flash_bm<DQK, DV, 64>(p, B, Hq, st);
}

// ------------------------------------------------------------------------------------------------ paged decode
constexpr int DEC_D = 128;     // head dim
constexpr int DEC_GROUP = 4;   // query heads per kv head (32 / 8)

struct DecodeAttnParams {
    const __nv_bfloat16* q;        // row b at q + b * q_stride: [n_q, D]
    long long q_stride;
    const __nv_bfloat16* k_cache;  // [slots, n_kv, D]
    const __nv_bfloat16* v_cache;
    const int* block_table;        // [B, max_pages]
    const int* cur_pos;            // [B] position of the current token; context = cur_pos + 1
    __nv_bfloat16* out;            // [B, n_q * D]
    float* part_o;                 // [B, n_kv, splits, GROUP, D]
    float* part_ml;                // [B, n_kv, splits, GROUP, 2]
    int* tickets;                  // [B, n_kv], zero-initialised, self-resetting
    int n_q, n_kv, page_size, max_pages, splits;
    float scale_log2;
    int early;                     // fetch cached K/V rows before the dependency wait
};

// ordered 16-byte load: volatile asm keeps it on its side of the dependency wait
__device__ __forceinline__ uint4 ld_early_v4(const void* ptr) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
    return r;
}

__device__ __forceinline__ unsigned long long attn_ld_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void attn_st_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// waits for a tagged word {value, tag != 0}, returns the value and frees the slot
__device__ __forceinline__ float attn_take(unsigned long long* addr, unsigned long long w) {
    for (uint32_t spin = 0; (w >> 32) == 0; ++spin) {
        if (spin > (1u << 22)) __trap();   // a split never published
        w = attn_ld_u64(addr);
    }
    attn_st_u64(addr, 0ull);
    return __uint_as_float(static_cast<uint32_t>(w));
}
__device__ __forceinline__ float attn_rescale(float m_old, float m_new) {
    return m_old == -INFINITY ? 0.0f : exp2f(m_old - m_new);
}

// TAGGED = false: splits are combined by the last CTA to take a ticket (fence + atomic + reload).
// TAGGED = true : partials travel as 64-bit {value, tag} words in an all-to-all over the splits of a kv head: each CTA
//                 owns a slice of the outputs and collects the other splits' words for it (no fence, no atomic, no
//                 serial collector); lane groups are pre-merged with shuffles.
template <bool TAGGED>
__global__ void __launch_bounds__(128)
decode_attn_kernel(const DecodeAttnParams p VITA_TRACE_PARAM) {
    __shared__ float s_m[16][DEC_GROUP];
    __shared__ float s_l[16][DEC_GROUP];
    __shared__ __align__(16) float s_o[TAGGED ? 4 : 16][DEC_GROUP][DEC_D];
    __shared__ float s_ml[16][DEC_GROUP][2];
    __shared__ int s_last;

    const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
#ifdef VITA_TRACE
    if (kvh != 0 || b != 0) trace = nullptr;
    if (threadIdx.x == 0 && split == 0) { VITA_STAMP(1); VITA_STAMP_SET(0, 2ull); }
#endif
    pdl_launch_dependents();
    if (!p.early) pdl_wait();
    // ---- before the dependency wait -------------------------------------------------------------------------------
    // cur_pos (first kernel of the step), the block table and every cached K/V row except the newest one are NOT
    // written by the kernel right before this one (the QKV projection of this layer), and the kernel before that has
    // completed by the time this grid could be launched (each chain kernel waits before it triggers).  So with
    // programmatic dependent launch the old keys are already in registers when the QKV kernel retires.
    const int ctx = p.cur_pos[b] + 1;
    const int newest = ctx - 1;
    const int per = (ctx + p.splits - 1) / p.splits;
    const int k_begin = split * per;
    const int k_end = min(ctx, k_begin + per);

    const int lg = threadIdx.x >> 3;  // lane group 0..15: one key at a time
    const int sl = threadIdx.x & 7;   // 16 dims per lane
    const int d0 = sl * 16;
    const int* bt = p.block_table + static_cast<long long>(b) * p.max_pages;

    constexpr int NPRE = 4;           // keys per lane group fetched up front (covers ctx <= splits * 64)
    uint4 pk[NPRE][2], pv[NPRE][2];
    long long pslot[NPRE];
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
        const int key = k_begin + j * 16 + lg;
        pslot[j] = -1;
        if (key < k_end) {
            const int page = bt[key / p.page_size];
            pslot[j] = static_cast<long long>(page) * p.page_size + key % p.page_size;
        }
    }
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
        const int key = k_begin + j * 16 + lg;
        if (pslot[j] >= 0 && key != newest) {
            const __nv_bfloat16* kp = p.k_cache + (pslot[j] * p.n_kv + kvh) * DEC_D + d0;
            const __nv_bfloat16* vp = p.v_cache + (pslot[j] * p.n_kv + kvh) * DEC_D + d0;
            pk[j][0] = ld_early_v4(kp); pk[j][1] = ld_early_v4(kp + 8);
            pv[j][0] = ld_early_v4(vp); pv[j][1] = ld_early_v4(vp + 8);
        } else {
            pk[j][0] = pk[j][1] = pv[j][0] = pv[j][1] = make_uint4(0, 0, 0, 0);
        }
    }
    pdl_wait();   // q and the newest K/V row come from the QKV kernel
    if (threadIdx.x == 0 && split == 0) VITA_STAMP(3);
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
        const int key = k_begin + j * 16 + lg;
        if (pslot[j] >= 0 && key == newest) {
            const __nv_bfloat16* kp = p.k_cache + (pslot[j] * p.n_kv + kvh) * DEC_D + d0;
            const __nv_bfloat16* vp = p.v_cache + (pslot[j] * p.n_kv + kvh) * DEC_D + d0;
            pk[j][0] = ld_early_v4(kp); pk[j][1] = ld_early_v4(kp + 8);
            pv[j][0] = ld_early_v4(vp); pv[j][1] = ld_early_v4(vp + 8);
        }
    }

    float q[DEC_GROUP][16];
#pragma unroll
    for (int h = 0; h < DEC_GROUP; ++h) {
        const uint4* qp = reinterpret_cast<const uint4*>(p.q + static_cast<long long>(b) * p.q_stride + (kvh * DEC_GROUP + h) * DEC_D + d0);
        const uint4 a = qp[0], c = qp[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { q[h][2 * i] = bf16_lo(w[i]); q[h][2 * i + 1] = bf16_hi(w[i]); }
    }
    float m[DEC_GROUP], l[DEC_GROUP], acc[DEC_GROUP][16];
#pragma unroll
    for (int h = 0; h < DEC_GROUP; ++h) {
        m[h] = -INFINITY;
        l[h] = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[h][i] = 0.0f;
    }
    // one key per lane group: scores for the 4 query heads of this kv head, online softmax, P.V
    auto consume = [&](const uint4& ka, const uint4& kc, const uint4& va, const uint4& vc, bool valid) {
        const uint32_t kw[8] = {ka.x, ka.y, ka.z, ka.w, kc.x, kc.y, kc.z, kc.w};
        const uint32_t vw[8] = {va.x, va.y, va.z, va.w, vc.x, vc.y, vc.z, vc.w};
        float kf[16], vf[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            kf[2 * i] = bf16_lo(kw[i]); kf[2 * i + 1] = bf16_hi(kw[i]);
            vf[2 * i] = bf16_lo(vw[i]); vf[2 * i + 1] = bf16_hi(vw[i]);
        }
#pragma unroll
        for (int h = 0; h < DEC_GROUP; ++h) {
            float dot = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i) dot += q[h][i] * kf[i];
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            if (valid) {
                const float sc = dot * p.scale_log2;
                const float m_new = fmaxf(m[h], sc);
                const float alpha = exp2f(m[h] - m_new);
                const float pr = exp2f(sc - m_new);
                m[h] = m_new;
                l[h] = l[h] * alpha + pr;
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[h][i] = acc[h][i] * alpha + pr * vf[i];
            }
        }
    };
    const int n_keys = max(k_end - k_begin, 0);
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
        if (j * 16 < n_keys)   // uniform across the CTA (shuffles inside)
            consume(pk[j][0], pk[j][1], pv[j][0], pv[j][1], pslot[j] >= 0);
    }
    // longer contexts: software pipeline, the K/V rows of the next key are in flight while the current one is reduced
    if (n_keys > NPRE * 16) {
        uint4 nk0, nk1, nv0, nv1;
        auto issue = [&](int key) {
            if (key < k_end) {
                const int page = bt[key / p.page_size];
                const long long slot = static_cast<long long>(page) * p.page_size + key % p.page_size;
                const uint4* kp = reinterpret_cast<const uint4*>(p.k_cache + (slot * p.n_kv + kvh) * DEC_D + d0);
                const uint4* vp = reinterpret_cast<const uint4*>(p.v_cache + (slot * p.n_kv + kvh) * DEC_D + d0);
                nk0 = kp[0]; nk1 = kp[1]; nv0 = vp[0]; nv1 = vp[1];
            } else {
                nk0 = nk1 = nv0 = nv1 = make_uint4(0, 0, 0, 0);
            }
        };
        issue(k_begin + NPRE * 16 + lg);
        for (int key0 = k_begin + NPRE * 16; key0 < k_end; key0 += 16) {  // trip count is uniform across the warp
            const int key = key0 + lg;
            const uint4 ka = nk0, kc = nk1, va = nv0, vc = nv1;
            issue(key + 16);
            consume(ka, kc, va, vc, key < k_end);
        }
    }
    if constexpr (TAGGED) {
        // ---- the 4 lane groups of each warp first (shuffles), then the 4 warps through shared memory
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
#pragma unroll
            for (int h = 0; h < DEC_GROUP; ++h) {
                const float mo = __shfl_xor_sync(0xffffffffu, m[h], off);
                const float lo = __shfl_xor_sync(0xffffffffu, l[h], off);
                const float mn = fmaxf(m[h], mo);
                const float wa = attn_rescale(m[h], mn), wb = attn_rescale(mo, mn);
                l[h] = l[h] * wa + lo * wb;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    acc[h][i] = acc[h][i] * wa + __shfl_xor_sync(0xffffffffu, acc[h][i], off) * wb;
                m[h] = mn;
            }
        }
        const int wid = threadIdx.x >> 5;
        if ((threadIdx.x & 31) < 8) {
#pragma unroll
            for (int h = 0; h < DEC_GROUP; ++h) {
                if (sl == 0) { s_m[wid][h] = m[h]; s_l[wid][h] = l[h]; }
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(&s_o[wid][h][d0 + i]) = make_float4(acc[h][i], acc[h][i + 1], acc[h][i + 2], acc[h][i + 3]);
            }
        }
        __syncthreads();
        const int d = threadIdx.x;  // one output dim per thread
        float mm[DEC_GROUP], ll[DEC_GROUP], oo[DEC_GROUP];
#pragma unroll
        for (int h = 0; h < DEC_GROUP; ++h) {
            mm[h] = fmaxf(fmaxf(s_m[0][h], s_m[1][h]), fmaxf(s_m[2][h], s_m[3][h]));
            ll[h] = 0.0f; oo[h] = 0.0f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float sc = attn_rescale(s_m[w][h], mm[h]);
                ll[h] += sc * s_l[w][h];
                oo[h] += sc * s_o[w][h][d];
            }
        }
        __syncthreads();   // every thread is done reading s_m / s_l / s_o
        // ---- all-to-all over the splits of this kv head: every CTA publishes its partials as tagged words and owns
        // a 1/splits slice of the 4 x 128 outputs, for which it collects the other splits' words (one round trip,
        // no fence, no ticket, no serial collector).  Each word has exactly one reader, which clears the tag.
        const int S = p.splits;                    // 4, 8 or 16 (checked on the host)
        const int P = DEC_GROUP * DEC_D / S;       // outputs owned by this CTA
        unsigned long long* words = reinterpret_cast<unsigned long long*>(p.part_o);
        const long long n_words_o = static_cast<long long>(gridDim.z) * p.n_kv * S * DEC_GROUP * DEC_D;
        const long long grp = static_cast<long long>(b) * p.n_kv + kvh;
        unsigned long long* o_w = words + grp * S * DEC_GROUP * DEC_D;                    // [publisher][head][dim]
        unsigned long long* ml_w = words + n_words_o + grp * S * S * DEC_GROUP * 2;      // [publisher][reader][head][2]
#ifdef VITA_TRACE
        if (threadIdx.x == 0 && trace) trace[9 + (split & 15)] = global_timer_ns();
#endif
        // publish
#pragma unroll
        for (int h = 0; h < DEC_GROUP; ++h) {
            s_o[0][h][d] = oo[h];   // own slice is read back from shared memory (s_o[0] is free again: all reads done)
            if ((h * DEC_D + d) / P != split)
                attn_st_u64(o_w + (split * DEC_GROUP + h) * DEC_D + d,
                            (1ull << 32) | static_cast<unsigned long long>(__float_as_uint(oo[h])));
        }
        const int t = threadIdx.x;
        const bool ml_thread = t < S * DEC_GROUP * 2;
        const int peer = t >> 3, mh = (t & 7) >> 1, which = t & 1;
        if (ml_thread) {
            float val = 0.0f;
#pragma unroll
            for (int h = 0; h < DEC_GROUP; ++h)
                if (h == mh) val = which ? ll[h] : mm[h];
            if (peer == split) s_ml[split][mh][which] = val;
            else attn_st_u64(ml_w + ((split * S + peer) * DEC_GROUP + mh) * 2 + which,
                             (1ull << 32) | static_cast<unsigned long long>(__float_as_uint(val)));
        }
        // collect: all loads first, then poll what has not arrived yet
        constexpr int MAXS = 16;
        unsigned long long* ml_addr = ml_w + ((peer * S + split) * DEC_GROUP + mh) * 2 + which;
        unsigned long long ml_word = 0;
        if (ml_thread && peer != split) ml_word = attn_ld_u64(ml_addr);
        const int flat = split * P + t;            // output handled by thread t < P
        const int oh = flat / DEC_D, od = flat % DEC_D;
        unsigned long long w[MAXS];
        float os[MAXS];
        if (t < P) {
#pragma unroll
            for (int sp = 0; sp < MAXS; ++sp)
                if (sp < S && sp != split) w[sp] = attn_ld_u64(o_w + (sp * DEC_GROUP + oh) * DEC_D + od);
        }
        if (ml_thread && peer != split) s_ml[peer][mh][which] = attn_take(ml_addr, ml_word);
        if (t < P) {
#pragma unroll
            for (int sp = 0; sp < MAXS; ++sp)
                os[sp] = (sp < S && sp != split) ? attn_take(o_w + (sp * DEC_GROUP + oh) * DEC_D + od, w[sp]) : 0.0f;
        }
        __syncthreads();
        if (threadIdx.x == 0) VITA_STAMP(27);
        if (t < P) {
            float mt = -INFINITY;
#pragma unroll
            for (int sp = 0; sp < MAXS; ++sp)
                if (sp < S) mt = fmaxf(mt, s_ml[sp][oh][0]);
            float lt = 0.0f, ot = 0.0f;
#pragma unroll
            for (int sp = 0; sp < MAXS; ++sp) {
                if (sp < S) {
                    const float sc = attn_rescale(s_ml[sp][oh][0], mt);
                    lt += sc * s_ml[sp][oh][1];
                    ot += sc * (sp == split ? s_o[0][oh][od] : os[sp]);
                }
            }
            p.out[(static_cast<long long>(b) * p.n_q + kvh * DEC_GROUP + oh) * DEC_D + od] =
                __float2bfloat16(lt > 0.0f ? ot / lt : 0.0f);
        }
        if (threadIdx.x == 0) VITA_STAMP(8);
        return;
    }
    // merge the 16 lane groups of this CTA
#pragma unroll
    for (int h = 0; h < DEC_GROUP; ++h) {
        if (sl == 0) { s_m[lg][h] = m[h]; s_l[lg][h] = l[h]; }
#pragma unroll
        for (int i = 0; i < 16; ++i) s_o[TAGGED ? 0 : lg][h][d0 + i] = acc[h][i];
    }
    __syncthreads();
    const int d = threadIdx.x;  // one output dim per thread
    const long long pbase = ((static_cast<long long>(b) * p.n_kv + kvh) * p.splits + split) * DEC_GROUP;
#pragma unroll
    for (int h = 0; h < DEC_GROUP; ++h) {
        float mm = -INFINITY;
#pragma unroll
        for (int gI = 0; gI < 16; ++gI) mm = fmaxf(mm, s_m[gI][h]);
        float ll = 0.0f, oo = 0.0f;
        if (mm > -INFINITY) {
#pragma unroll
            for (int gI = 0; gI < 16; ++gI) {
                const float w = exp2f(s_m[gI][h] - mm);
                ll += w * s_l[gI][h];
                oo += w * s_o[TAGGED ? 0 : gI][h][d];
            }
        }
        p.part_o[(pbase + h) * DEC_D + d] = oo;
        if (d == 0) { p.part_ml[(pbase + h) * 2] = mm; p.part_ml[(pbase + h) * 2 + 1] = ll; }
    }
#ifdef VITA_TRACE
    if (threadIdx.x == 0 && trace) trace[9 + (split & 15)] = global_timer_ns();
#endif
    // last CTA of this (batch, kv head) merges the splits
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(&p.tickets[b * p.n_kv + kvh], 1);
        s_last = (t == p.splits - 1);
        if (s_last) p.tickets[b * p.n_kv + kvh] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) VITA_STAMP(25);
    __threadfence();
    const long long sbase = (static_cast<long long>(b) * p.n_kv + kvh) * p.splits * DEC_GROUP;
    constexpr int MAXS = 16;   // splits <= 16 (checked on the host); all loads of a head are issued before use
#pragma unroll
    for (int h = 0; h < DEC_GROUP; ++h) {
        float ms[MAXS], ls[MAXS], os[MAXS];
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
            if (s < p.splits) {
                ms[s] = __ldcg(&p.part_ml[(sbase + s * DEC_GROUP + h) * 2]);
                ls[s] = __ldcg(&p.part_ml[(sbase + s * DEC_GROUP + h) * 2 + 1]);
                os[s] = __ldcg(&p.part_o[(sbase + s * DEC_GROUP + h) * DEC_D + d]);
            } else {
                ms[s] = -INFINITY; ls[s] = 0.0f; os[s] = 0.0f;
            }
        }
        // consumed: hand the workspace back all-zero (the tagged variant shares it and reads a non-zero upper half
        // of an 8-byte word as "published"; each element below was read by this thread only)
#pragma unroll
        for (int s = 0; s < MAXS; ++s)
            if (s < p.splits) __stcg(&p.part_o[(sbase + s * DEC_GROUP + h) * DEC_D + d], 0.0f);
        float mm = -INFINITY;
#pragma unroll
        for (int s = 0; s < MAXS; ++s) mm = fmaxf(mm, ms[s]);
        float ll = 0.0f, oo = 0.0f;
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
            const float w = (ms[s] == -INFINITY) ? 0.0f : exp2f(ms[s] - mm);
            ll += w * ls[s];
            oo += w * os[s];
        }
        p.out[(static_cast<long long>(b) * p.n_q + kvh * DEC_GROUP + h) * DEC_D + d] =
            __float2bfloat16(ll > 0.0f ? oo / ll : 0.0f);
    }
    __syncthreads();   // every thread has read the (m, l) pairs
    if (d < p.splits * DEC_GROUP * 2) __stcg(&p.part_ml[sbase * 2 + d], 0.0f);
    if (threadIdx.x == 0) VITA_STAMP(8);
}

}  // namespace vita

using namespace vita;

extern "C" int vita_attention_fwd(const void* q, const void* k, const void* v, void* o, const int64_t* q_strides,
                                  const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides,
                                  int64_t B, int64_t n_q_heads, int64_t n_kv_heads, int64_t Sq, int64_t Skv,
                                  int64_t d_qk, int64_t d_v, const int32_t* kv_lens, int causal, float scale,
                                  void* stream) {
    VITA_REQUIRE(n_kv_heads > 0 && n_q_heads % n_kv_heads == 0, "n_q_heads must be a multiple of n_kv_heads");
    VITA_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), "q/k/v/o must be 16-byte aligned");
    for (int i = 0; i < 3; ++i)
        VITA_REQUIRE(q_strides[i] % 8 == 0 && k_strides[i] % 8 == 0 && v_strides[i] % 8 == 0 && o_strides[i] % 2 == 0,
                     "strides must keep 16-byte row alignment");
    if (B == 0 || Sq == 0) return VITA_OK;
    AttnParams p{};
    p.q = BF16C(q); p.k = BF16C(k); p.v = BF16C(v); p.o = static_cast<__nv_bfloat16*>(o);
    p.q_bs = q_strides[0]; p.q_ts = q_strides[1]; p.q_hs = q_strides[2];
    p.k_bs = k_strides[0]; p.k_ts = k_strides[1]; p.k_hs = k_strides[2];
    p.v_bs = v_strides[0]; p.v_ts = v_strides[1]; p.v_hs = v_strides[2];
    p.o_bs = o_strides[0]; p.o_ts = o_strides[1]; p.o_hs = o_strides[2];
    p.group = static_cast<int>(n_q_heads / n_kv_heads);
    p.Sq = static_cast<int>(Sq);
    p.Skv = static_cast<int>(Skv);
    p.kv_lens = kv_lens;
    p.causal = causal;
    p.scale_log2 = scale * 1.4426950408889634f;
    auto st = static_cast<cudaStream_t>(stream);
    if (d_qk == 128 && d_v == 128) return launch_flash<128, 128>(p, (int)B, (int)n_q_heads, st);
    if (d_qk == 64 && d_v == 64) return launch_flash<64, 64>(p, (int)B, (int)n_q_heads, st);
    if (d_qk == 128 && d_v == 64) return launch_flash<128, 64>(p, (int)B, (int)n_q_heads, st);
    set_last_error("vita_attention_fwd: unsupported head dims (supported: 128/128, 64/64, 128/64)");
    return VITA_ERR_INVALID;
}

extern "C" int64_t vita_decode_attention_workspace_bytes(int64_t B, int64_t n_kv_heads, int64_t splits) {
    // 8-byte words {value, tag} for the split partials; the ticket variant uses the front of the same buffer as plain
    // floats.  Invariant kept by both variants: the workspace is all-zero between launches (tags cleared by their
    // readers, floats zeroed by the merging CTA, tickets reset), so calls may change B / splits / variant freely.
    const int64_t n = B * n_kv_heads * splits * DEC_GROUP;
    return n * DEC_D * 8 + B * n_kv_heads * splits * splits * DEC_GROUP * 2 * 8 + ((B * n_kv_heads * 4 + 255) / 256) * 256 + 256;
}

extern "C" int vita_decode_attention(const void* q, const void* k_cache, const void* v_cache,
                                     const int32_t* block_table, const int32_t* cur_pos, void* out, void* workspace,
                                     int64_t B, int64_t n_q_heads, int64_t n_kv_heads, int64_t head_dim,
                                     int64_t page_size, int64_t max_pages, int64_t splits, float scale, int64_t q_stride,
                                     void* stream) {
    VITA_REQUIRE(head_dim == DEC_D && n_q_heads == n_kv_heads * DEC_GROUP, "decode attention: need D=128, GQA group 4");
    VITA_REQUIRE(splits >= 1 && splits <= 16 && workspace != nullptr, "1 <= splits <= 16 and a workspace are required");
    if (B == 0) return VITA_OK;
    DecodeAttnParams p{};
    p.q = BF16C(q); p.q_stride = q_stride > 0 ? q_stride : n_q_heads * head_dim;
    p.k_cache = BF16C(k_cache); p.v_cache = BF16C(v_cache);
    p.block_table = block_table; p.cur_pos = cur_pos; p.out = static_cast<__nv_bfloat16*>(out);
    const int64_t n = B * n_kv_heads * splits * DEC_GROUP;
    // tickets first (must be zero-initialised once by the caller; the kernel resets them)
    p.tickets = static_cast<int*>(workspace);
    char* w = static_cast<char*>(workspace) + ((B * n_kv_heads * 4 + 255) / 256) * 256;
    p.part_o = reinterpret_cast<float*>(w);
    p.part_ml = reinterpret_cast<float*>(w + n * DEC_D * 4);
    p.n_q = (int)n_q_heads; p.n_kv = (int)n_kv_heads; p.page_size = (int)page_size; p.max_pages = (int)max_pages;
    p.splits = (int)splits;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.early = option("attn_early");
    dim3 grid((unsigned)splits, (unsigned)n_kv_heads, (unsigned)B);
    // the all-to-all protocol needs the split count to divide the 4 x 128 outputs into <= 128-wide slices
    // ... and every CTA of a kv head polls its peers, so all of them have to be resident at once: keep that to grids
    // that fit the machine (two 128-thread CTAs per SM), larger batches take the ticket variant
    const bool tagged = option("attn_tagged") && (splits == 4 || splits == 8 || splits == 16) &&
                        splits * n_kv_heads * B <= 2ll * (num_sms() > 0 ? num_sms() : 148);
    {
        static int carveout = -1;
        const int want = option("smem_carveout_max") ? cudaSharedmemCarveoutMaxShared : cudaSharedmemCarveoutDefault;
        if (carveout != want) {
            (void)cudaFuncSetAttribute(decode_attn_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, want);
            (void)cudaFuncSetAttribute(decode_attn_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, want);
            carveout = want;
        }
    }
    cudaError_t e = tagged
        ? launch_chain(decode_attn_kernel<true>, grid, dim3(128), 0, static_cast<cudaStream_t>(stream), p VITA_TRACE_ARG)
        : launch_chain(decode_attn_kernel<false>, grid, dim3(128), 0, static_cast<cudaStream_t>(stream), p VITA_TRACE_ARG);
    if (e != cudaSuccess) return check_cuda(e, "decode_attn_kernel");
    return check_launch("decode_attn_kernel");
}
