// bf16 "TN" GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . B[N,K]^T), fp32 accumulation in TMEM.
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0   : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1   : MMA issuer    (one thread, tcgen05.mma cta_group::1 kind::f16, 128 x BLOCK_N x 16 per instruction)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      (tcgen05.ld 32x32b -> registers -> fused epilogue -> 64 B vector stores)
// The accumulator is double buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  The same kernel serves the plain linears (ViT / Whale / projector / Mixtral qkv,o) and the
// grouped expert GEMMs of the MoE (ragged row groups given by a device-side offsets array, no host sync).
//
// Reference call sites this replaces (all nn.Linear / F.linear -> cuBLAS in the reference):
//   vita/model/multimodal_encoder/internvit/modeling_intern_vit.py:180,192,214,216
//   vita/model/multimodal_encoder/whale/module/layer/attention.py:371-373,381,419 ; :145-147
//   vita/model/multimodal_projector/builder.py:164-168
//   transformers MixtralAttention q/k/v/o_proj, MixtralExperts.forward (gate_up_proj / down_proj)
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace vita {

struct GemmArgs {
    int M;                      // rows of A / C (total over groups)
    int N;                      // output columns (per group)
    int K;
    int num_groups;             // >= 1
    const int* group_offsets;   // device [num_groups + 1] row offsets, or nullptr (single group = all rows)
    const int* group_counts;    // or: group g occupies rows [g * group_stride, g * group_stride + group_counts[g])
    int group_stride;           //     ("slot" layout written by vita_moe_route_scatter: a fixed capacity per expert)
    int est_rows;               // host-side estimate of the occupied rows (tile-shape heuristics); 0 = M
    __nv_bfloat16* C;
    long long ldc;
    const __nv_bfloat16* bias;  // [num_groups, N] or nullptr
    const __nv_bfloat16* colscale;  // [N] or nullptr   (InternViT layer-scale ls1/ls2)
    const float* rowscale;          // [M] or nullptr   (MoE routing weight)
    const __nv_bfloat16* residual;  // [M, ldr] or nullptr
    long long ldr;
    int act;                    // VITA_ACT_*
    // expert-parallel scatter epilogue (down projection): output row r goes to the rank that owns its token,
    //   dst = peer_out[t / chunk] + ((t % chunk) * 2 + k) * ldc   with (t, k) = (row_assign[r] >> 1, row_assign[r] & 1)
    const int* row_assign;
    __nv_bfloat16* const* peer_out;
    int chunk;
    int tail_split;             // cut the tiles of the last partial wave along N (see Sched)
    // fused RoPE + paged KV append epilogue of the qkv projection (N = (n_q + 2 n_kv) * 128; MixtralAttention.forward
    // modeling_mixtral.py:312-340 + the cache update): column block h of 128 is head h; q and k heads are rotated with
    // the row's position, k and v heads are also written to their cache slot.  nullptr = plain epilogue.
    const float* rope_cos_sin;  // fp32 [max_pos, 2, 64]
    const int* rope_pos;        // [M] position of each row
    const int* rope_slot;       // [M] cache slot of each row, or nullptr (no cache write)
    __nv_bfloat16* k_cache;
    __nv_bfloat16* v_cache;
    int n_q, n_kv;
};

__device__ __forceinline__ void group_rows(const GemmArgs& a, int g, int& r0, int& r1) {
    if (a.group_counts) { r0 = g * a.group_stride; r1 = r0 + a.group_counts[g]; }
    else if (a.group_offsets) { r0 = a.group_offsets[g]; r1 = a.group_offsets[g + 1]; }
    else { r0 = 0; r1 = a.M; }
}

struct Tile {
    int group, m0, m_end, n0;
    int bn;   // MMA N of this work item: BLOCK_N for a whole tile, BLOCK_N / split for a piece of a tail tile
};

// Work list of one launch.  Tiles are ordered group-major, n-block major, m-block minor (CTAs that run together share
// the weight tile in L2) and dealt round-robin to the persistent CTAs.  The tiles of the last, partly filled wave are
// cut along N into `split` pieces (2 or 4) when that still fits one wave: a 512-tile problem on 148 SMs then costs
// 3 + 1/2 waves instead of 4, a 1040-tile one 7 + 1/4 instead of 8.  (The fused gate|up variant is not cut: its N
// tile is two half tiles and it runs ~50 waves.)
struct Sched {
    int total;      // whole tiles
    int full_end;   // tiles [0, full_end) are processed whole
    int split;      // the others in `split` pieces each
    __device__ __forceinline__ int items() const { return full_end + (total - full_end) * split; }
};

template <int BLOCK_N, int BN_OUT, bool SILU, int TILE_M>
__device__ __forceinline__ Sched make_sched(const GemmArgs& a) {
    const int num_n = (a.N + BN_OUT - 1) / BN_OUT;
    int total = 0;
    for (int g = 0; g < a.num_groups; ++g) {
        int r0, r1;
        group_rows(a, g, r0, r1);
        total += ((r1 - r0 + TILE_M - 1) / TILE_M) * num_n;
    }
    Sched s{total, total, 1};
    if (!SILU && a.tail_split) {
        const int G = static_cast<int>(gridDim.x);
        const int R = total % G;
        constexpr int MAX_SPLIT = BLOCK_N / 32 < 4 ? BLOCK_N / 32 : 4;   // pieces of at least 32 columns, at most 4
        const int max_split = a.rope_cos_sin ? BLOCK_N / 128 : MAX_SPLIT;   // RoPE pairs column j with j + 64 of a head
        int sp = 1;
        while (R > 0 && sp * 2 <= max_split && sp * 2 * R <= G) sp *= 2;
        if (sp > 1) { s.full_end = total - R; s.split = sp; }
    }
    return s;
}

template <int BLOCK_N, int BN_OUT, int TILE_M>
__device__ __forceinline__ bool tile_at(const GemmArgs& a, const Sched& sc, int item, Tile& t) {
    if (item >= sc.items()) return false;
    int tile_idx = item, part = 0;
    t.bn = BLOCK_N;
    if (item >= sc.full_end) {
        const int j = item - sc.full_end;
        tile_idx = sc.full_end + j / sc.split;
        part = j % sc.split;
        t.bn = BLOCK_N / sc.split;
    }
    const int num_n = (a.N + BN_OUT - 1) / BN_OUT;
    int base = 0;
    for (int g = 0; g < a.num_groups; ++g) {
        int r0, r1;
        group_rows(a, g, r0, r1);
        const int mt = (r1 - r0 + TILE_M - 1) / TILE_M;
        const int nt = mt * num_n;
        if (tile_idx < base + nt) {
            const int local = tile_idx - base;
            t.group = g;
            t.n0 = (local / mt) * BN_OUT + part * t.bn;
            t.m0 = r0 + (local % mt) * TILE_M;
            t.m_end = r1;
            return true;
        }
        base += nt;
    }
    return false;
}

// MT = 1: one 128-row tile per work item, accumulator double buffered (the epilogue of a tile overlaps the next
// tile's MMAs).  MT = 2: two 128-row tiles share every B stage (one pass over the weights serves up to 256 rows of a
// group; the whole TMEM holds the two accumulators, so the epilogue is not overlapped) -- for grouped GEMMs whose
// groups hold 100-250 rows, where the second row tile of a group would otherwise stream the same weights again.
template <int BLOCK_N, bool SILU, int STAGES, int MT>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmBs /* B in boxes of 32 rows: pieces of tail tiles */,
                    const GemmArgs args) {
    constexpr int BLOCK_M = 128, BLOCK_K = 64;
    constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr int BN_OUT = SILU ? BLOCK_N / 2 : BLOCK_N;
    constexpr int TILE_M = BLOCK_M * MT;
    constexpr int A_STAGE = MT * A_BYTES;
    constexpr int NACC = MT == 1 ? 2 : 1;                 // accumulator stages
    constexpr int ACC_COLS = MT * BLOCK_N;                // TMEM columns of one stage
    constexpr uint32_t TMEM_COLS = NACC * ACC_COLS;
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (args.K + BLOCK_K - 1) / BLOCK_K;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (!SILU) tma_prefetch_desc(&tmBs);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            Tile t;
            const Sched sc = make_sched<BLOCK_N, BN_OUT, SILU, TILE_M>(args);
            for (int tile = blockIdx.x; tile_at<BLOCK_N, BN_OUT, TILE_M>(args, sc, tile, t); tile += gridDim.x) {
                const int n_m = (MT == 2 && t.m_end - t.m0 > BLOCK_M) ? 2 : 1;   // row tiles that hold valid rows
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    mbar_arrive_expect_tx(&full_bar[stage], n_m * A_BYTES + t.bn * (BLOCK_K * 2));
                    for (int mi = 0; mi < n_m; ++mi)
                        tma_load_2d(sA + stage * A_STAGE + mi * A_BYTES, &tmA, &full_bar[stage], kb * BLOCK_K,
                                    t.m0 + mi * BLOCK_M);
                    if (!SILU && t.bn != BLOCK_N) {
                        for (int r = 0; r < t.bn; r += 32)
                            tma_load_3d(sB + stage * B_BYTES + r * (BLOCK_K * 2), &tmBs, &full_bar[stage], kb * BLOCK_K,
                                        t.n0 + r, t.group);
                    } else if constexpr (SILU) {
                        // gate rows [n0, n0+128) and up rows [N + n0, N + n0 + 128) of the fused [2N, K] weight
                        tma_load_3d(sB + stage * B_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, t.n0, t.group);
                        tma_load_3d(sB + stage * B_BYTES + B_BYTES / 2, &tmB, &full_bar[stage], kb * BLOCK_K,
                                    args.N + t.n0, t.group);
                    } else {
                        tma_load_3d(sB + stage * B_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, t.n0, t.group);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------ MMA issuer
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
            Tile t;
            const Sched sc = make_sched<BLOCK_N, BN_OUT, SILU, TILE_M>(args);
            for (int tile = blockIdx.x; tile_at<BLOCK_N, BN_OUT, TILE_M>(args, sc, tile, t); tile += gridDim.x) {
                const uint32_t idesc = umma_idesc_bf16(BLOCK_M, t.bn);
                const int n_m = (MT == 2 && t.m_end - t.m0 > BLOCK_M) ? 2 : 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
                    const uint64_t db = umma_desc_k_sw128(sB_addr + stage * B_BYTES);
                    for (int mi = 0; mi < n_m; ++mi) {
                        const uint64_t da = umma_desc_k_sw128(sA_addr + stage * A_STAGE + mi * A_BYTES);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k) {
                            // advance 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in 16 B units
                            tc_mma_bf16(d_tmem + mi * BLOCK_N, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) tc_commit(&tmem_full[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue
        const int quad = warp - 4;  // == warp % 4: the TMEM lane quadrant this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        Tile t;
        const Sched sc = make_sched<BLOCK_N, BN_OUT, SILU, TILE_M>(args);
        for (int tile = blockIdx.x; tile_at<BLOCK_N, BN_OUT, TILE_M>(args, sc, tile, t); tile += gridDim.x) {
            const int n_chunks = (SILU ? BN_OUT : t.bn) / 32;
            const int n_m = (MT == 2 && t.m_end - t.m0 > BLOCK_M) ? 2 : 1;
            mbar_wait(&tmem_full[acc], acc_phase, 4);
            tc_fence_after();
          for (int mi = 0; mi < n_m; ++mi) {
            const int grow = t.m0 + mi * BLOCK_M + quad * 32 + lane;
            const bool valid = grow < t.m_end;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_COLS + mi * BLOCK_N;
            const float rs = (args.rowscale && valid) ? args.rowscale[grow] : 1.0f;
            const __nv_bfloat16* bias = args.bias ? args.bias + static_cast<long long>(t.group) * args.N : nullptr;
            __nv_bfloat16* crow = args.C + static_cast<long long>(grow) * args.ldc;
            if (args.peer_out != nullptr && valid) {
                const int a = args.row_assign[grow], tk = a >> 1;
                const int owner = tk / args.chunk;
                crow = args.peer_out[owner] + (static_cast<long long>(tk - owner * args.chunk) * 2 + (a & 1)) * args.ldc;
            }
            const __nv_bfloat16* rrow = args.residual ? args.residual + static_cast<long long>(grow) * args.ldr : nullptr;
            if constexpr (!SILU && MT == 1) {
                if (args.rope_cos_sin != nullptr) {
                    // ---- qkv projection: RoPE on the bf16-rounded projections, K / V rows appended to the paged cache
                    const int pos = valid ? args.rope_pos[grow] : 0;
                    const long long slot = (valid && args.rope_slot) ? args.rope_slot[grow] : -1;
                    const float4* cs = reinterpret_cast<const float4*>(args.rope_cos_sin + static_cast<long long>(pos) * 128);
#pragma unroll 1
                    for (int hh = 0; hh < t.bn / 128; ++hh) {
                        const int head = (t.n0 >> 7) + hh;
                        if (head * 128 >= args.N) break;   // warp-uniform
                        const bool rot = head < args.n_q + args.n_kv;
                        __nv_bfloat16* cdst = nullptr;
                        if (slot >= 0 && head >= args.n_q) {
                            const bool is_k = head < args.n_q + args.n_kv;
                            const int kvh = is_k ? head - args.n_q : head - args.n_q - args.n_kv;
                            cdst = (is_k ? args.k_cache : args.v_cache) + (slot * args.n_kv + kvh) * 128;
                        }
#pragma unroll 1
                        for (int half = 0; half < 2; ++half) {   // columns [32 half, 32 half + 32) and their partners + 64
                            uint32_t a[32], b[32];
                            tmem_ld_32x32(taddr + hh * 128 + half * 32, a);
                            tmem_ld_32x32(taddr + hh * 128 + 64 + half * 32, b);
                            tmem_ld_wait();
                            if (!valid) continue;
                            uint32_t oa[16], ob[16];
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                float4 cc = make_float4(1.f, 1.f, 1.f, 1.f), ss = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (rot) { cc = __ldg(cs + half * 8 + q); ss = __ldg(cs + 16 + half * 8 + q); }
                                const float c4[4] = {cc.x, cc.y, cc.z, cc.w}, s4[4] = {ss.x, ss.y, ss.z, ss.w};
                                float r1[4], r2[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    // projections are rounded to bf16 first, as the stand-alone pass sees them
                                    const float x1 = __bfloat162float(__float2bfloat16(__uint_as_float(a[q * 4 + e])));
                                    const float x2 = __bfloat162float(__float2bfloat16(__uint_as_float(b[q * 4 + e])));
                                    if (rot) rope_rotate(x1, x2, c4[e], s4[e], r1[e], r2[e]);
                                    else { r1[e] = x1; r2[e] = x2; }
                                }
                                oa[q * 2] = pack_bf16(r1[0], r1[1]); oa[q * 2 + 1] = pack_bf16(r1[2], r1[3]);
                                ob[q * 2] = pack_bf16(r2[0], r2[1]); ob[q * 2 + 1] = pack_bf16(r2[2], r2[3]);
                            }
                            uint4* d1 = reinterpret_cast<uint4*>(crow + head * 128 + half * 32);
                            uint4* d2 = reinterpret_cast<uint4*>(crow + head * 128 + 64 + half * 32);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                d1[q] = make_uint4(oa[q * 4], oa[q * 4 + 1], oa[q * 4 + 2], oa[q * 4 + 3]);
                                d2[q] = make_uint4(ob[q * 4], ob[q * 4 + 1], ob[q * 4 + 2], ob[q * 4 + 3]);
                            }
                            if (cdst != nullptr) {
                                uint4* k1 = reinterpret_cast<uint4*>(cdst + half * 32);
                                uint4* k2 = reinterpret_cast<uint4*>(cdst + 64 + half * 32);
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    k1[q] = make_uint4(oa[q * 4], oa[q * 4 + 1], oa[q * 4 + 2], oa[q * 4 + 3]);
                                    k2[q] = make_uint4(ob[q * 4], ob[q * 4 + 1], ob[q * 4 + 2], ob[q * 4 + 3]);
                                }
                            }
                        }
                    }
                    continue;   // (MT == 1: leaves the one-iteration row-tile loop; the accumulator is released below)
                }
            }
#pragma unroll 1
            for (int c = 0; c < n_chunks; ++c) {
                const int col0 = t.n0 + c * 32;
                if (col0 >= args.N) break;  // warp-uniform
                uint32_t v[32];
                tmem_ld_32x32(taddr + c * 32, v);
                if constexpr (SILU) {
                    uint32_t u[32];
                    tmem_ld_32x32(taddr + BN_OUT + c * 32, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        v[j] = __float_as_uint(silu(__uint_as_float(v[j])) * __uint_as_float(u[j]));
                } else {
                    tmem_ld_wait();
                }
                const bool full_chunk = (col0 + 32 <= args.N);
                if (full_chunk) {
                    // column-wise terms are identical for all lanes: broadcast loads
                    if (bias) {
                        const uint4* bp = reinterpret_cast<const uint4*>(bias + col0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 b = __ldg(bp + q);
                            const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                v[q * 8 + e * 2] = __float_as_uint(__uint_as_float(v[q * 8 + e * 2]) + bf16_lo(w[e]));
                                v[q * 8 + e * 2 + 1] =
                                    __float_as_uint(__uint_as_float(v[q * 8 + e * 2 + 1]) + bf16_hi(w[e]));
                            }
                        }
                    }
                    if (args.act == VITA_ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(gelu_erf(__uint_as_float(v[j])));
                    } else if (args.act == VITA_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.0f));
                    }
                    if (args.colscale) {
                        const uint4* sp = reinterpret_cast<const uint4*>(args.colscale + col0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 b = __ldg(sp + q);
                            const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                v[q * 8 + e * 2] = __float_as_uint(__uint_as_float(v[q * 8 + e * 2]) * bf16_lo(w[e]));
                                v[q * 8 + e * 2 + 1] =
                                    __float_as_uint(__uint_as_float(v[q * 8 + e * 2 + 1]) * bf16_hi(w[e]));
                            }
                        }
                    }
                    if (valid) {
                        if (args.rowscale) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * rs);
                        }
                        if (rrow) {
                            const uint4* rp = reinterpret_cast<const uint4*>(rrow + col0);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 b = rp[q];
                                const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    v[q * 8 + e * 2] =
                                        __float_as_uint(__uint_as_float(v[q * 8 + e * 2]) + bf16_lo(w[e]));
                                    v[q * 8 + e * 2 + 1] =
                                        __float_as_uint(__uint_as_float(v[q * 8 + e * 2 + 1]) + bf16_hi(w[e]));
                                }
                            }
                        }
                        uint4* cp = reinterpret_cast<uint4*>(crow + col0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 o;
                            o.x = pack_bf16(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
                            o.y = pack_bf16(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
                            o.z = pack_bf16(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
                            o.w = pack_bf16(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
                            cp[q] = o;
                        }
                    }
                } else if (valid) {
                    // ragged N tail: scalar path
                    for (int j = 0; j < 32; ++j) {
                        const int col = col0 + j;
                        if (col >= args.N) break;
                        float x = __uint_as_float(v[j]);
                        if (bias) x += __bfloat162float(bias[col]);
                        if (args.act == VITA_ACT_GELU) x = gelu_erf(x);
                        else if (args.act == VITA_ACT_RELU) x = fmaxf(x, 0.0f);
                        if (args.colscale) x *= __bfloat162float(args.colscale[col]);
                        x *= rs;
                        if (rrow) x += __bfloat162float(rrow[col]);
                        crow[col] = __float2bfloat16(x);
                    }
                }
            }
          }   // row tiles
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int BLOCK_N, bool SILU, int STAGES, int MT = 1>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBs, const GemmArgs& args,
                       int max_tiles, cudaStream_t stream) {
    constexpr int smem_bytes = STAGES * (MT * 128 * 64 * 2 + BLOCK_N * 64 * 2) + 1024 + 256;
    static bool configured = false;
    auto kern = gemm_bf16_tn_kernel<BLOCK_N, SILU, STAGES, MT>;
    if (!configured) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes),
                            "cudaFuncSetAttribute(gemm smem)");
        if (rc) return rc;
        configured = true;
    }
    int grid = num_sms();
    // with the tail split a problem of fewer tiles than SMs still spreads over the whole machine
    const long long items = static_cast<long long>(max_tiles) * ((!SILU && args.tail_split) ? (BLOCK_N / 32 < 4 ? BLOCK_N / 32 : 4) : 1);
    if (items < grid) grid = static_cast<int>(items);
    if (grid < 1) grid = 1;
    kern<<<grid, 256, smem_bytes, stream>>>(tmA, tmB, tmBs, args);
    return check_launch("gemm_bf16_tn_kernel");
}

// a_rows: number of valid rows in A (TMA zero-fills beyond); b: [num_groups, b_rows, K] contiguous.
static int gemm_dispatch(const void* A, long long lda, int a_rows, const void* B, int b_rows, const GemmArgs& args,
                         bool silu, cudaStream_t stream) {
    VITA_REQUIRE(args.K > 0 && args.N > 0 && args.M >= 0, "bad shape");
    VITA_REQUIRE(args.K % 8 == 0 && lda % 8 == 0, "K and lda must be multiples of 8 (16-byte TMA strides)");
    VITA_REQUIRE(aligned16(A) && aligned16(B) && aligned16(args.C), "A, B, C must be 16-byte aligned");
    VITA_REQUIRE(args.ldc % 8 == 0, "ldc must be a multiple of 8");
    VITA_REQUIRE(!args.residual || (args.ldr % 8 == 0 && aligned16(args.residual)), "residual alignment");
    VITA_REQUIRE(!args.bias || aligned16(args.bias), "bias alignment");
    VITA_REQUIRE(!args.bias || args.num_groups == 1 || args.N % 8 == 0, "grouped bias needs N % 8 == 0");
    VITA_REQUIRE(!args.colscale || aligned16(args.colscale), "colscale alignment");
    if (args.M == 0) return VITA_OK;

    const int n_sms = num_sms();
    const int occ_rows = args.est_rows > 0 ? args.est_rows : args.M;   // rows that carry work (slot layout: M is the capacity)
    const long long m_tiles_ub = (occ_rows / 128) + args.num_groups;  // valid for any ragged split of the occupied rows
    // Tile-shape choice.  Small-M problems stream each weight tile once and are HBM-bound, so what matters is how
    // evenly the tiles fill the 148 SMs (wave quantisation); large-M problems want the 256-wide tile (smem operand
    // bandwidth: 96 B/cycle/SM instead of 128).  Estimate the tile count and take the widest tile whose last wave
    // is not mostly idle.
    const long long m_tiles_est = (occ_rows + 127) / 128 + (args.num_groups > 1 ? args.num_groups / 2 : 0);
    const char* ts_env = getenv("VITA_B200_GEMM_TAIL_SPLIT");   // tuning aid: 0 switches the tail split off
    const bool tail_split = !silu && !(ts_env && atoi(ts_env) == 0);
    // Estimated duration in units of one 128 x 256 tile on one SM.  Measured on B200 (profiles/r02_gemm_shapes.txt):
    // a 128-wide tile costs 0.76 of a 256-wide one when compute-bound (its A operand is read twice as often per flop),
    // half when the weights are streamed once from HBM (few row tiles); the pieces of a cut tail tile likewise
    // 0.75 / 0.6 of the whole for 1/2 and 1/4 of the width when compute-bound, 1/2 and 1/4 when HBM-bound.
    const bool hbm_bound = m_tiles_est <= 2;
    auto est_time = [&](int bn) {
        const long long tiles = m_tiles_est * ((args.N + bn - 1) / bn);
        const long long whole = tiles / n_sms, rest = tiles % n_sms;
        double waves = static_cast<double>(whole);
        if (rest > 0) {
            int sp = 1;
            while (tail_split && sp < 4 && sp * 2 * rest <= n_sms) sp *= 2;
            waves += hbm_bound ? 1.0 / sp : (sp == 1 ? 1.0 : (sp == 2 ? 0.75 : 0.6));
        }
        return waves * (bn == 256 ? 1.0 : (hbm_bound ? 0.5 : 0.76));
    };
    int block_n;
    const char* force = getenv("VITA_B200_GEMM_BN");   // tuning aid: force the tile width (128 / 256)
    if (force && (atoi(force) == 128 || atoi(force) == 256)) {
        block_n = atoi(force);
        if (!silu && args.N < 256) block_n = 128;
    } else if (silu) {
        // the 64-wide variant halves the weight tile but re-reads the activation tile twice as often from L2:
        // measured slower (474 vs 388 us per layer at S=506), so it is only used when forced
        block_n = 256;
    } else if (args.rope_cos_sin != nullptr) {
        block_n = 256;   // two whole heads per tile
    } else if (args.N < 256) {
        block_n = 128;
    } else {
        block_n = (est_time(256) <= est_time(128) * 1.02) ? 256 : 128;
    }
    const int bn_out = silu ? block_n / 2 : block_n;
    const long long max_tiles_ll = m_tiles_ub * ((args.N + bn_out - 1) / bn_out);
    const int max_tiles = max_tiles_ll > (1 << 30) ? (1 << 30) : static_cast<int>(max_tiles_ll);

    GemmArgs a2 = args;
    a2.tail_split = tail_split ? 1 : 0;
    CUtensorMap tmA, tmB, tmBs;
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(args.K), static_cast<uint64_t>(a_rows)};
        const uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
        const uint32_t box[2] = {64, 128};
        int rc = make_tensor_map_bf16(&tmA, A, 2, dims, strides, box, true);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {static_cast<uint64_t>(args.K), static_cast<uint64_t>(b_rows),
                                  static_cast<uint64_t>(args.num_groups)};
        const uint64_t strides[2] = {static_cast<uint64_t>(args.K) * 2,
                                     static_cast<uint64_t>(args.K) * 2 * static_cast<uint64_t>(b_rows)};
        const uint32_t box[3] = {64, static_cast<uint32_t>(silu ? block_n / 2 : block_n), 1};
        int rc = make_tensor_map_bf16(&tmB, B, 3, dims, strides, box, true);
        if (rc) return rc;
    }
    tmBs = tmB;
    if (tail_split) {
        const uint64_t dims[3] = {static_cast<uint64_t>(args.K), static_cast<uint64_t>(b_rows),
                                  static_cast<uint64_t>(args.num_groups)};
        const uint64_t strides[2] = {static_cast<uint64_t>(args.K) * 2,
                                     static_cast<uint64_t>(args.K) * 2 * static_cast<uint64_t>(b_rows)};
        const uint32_t box[3] = {64, 32, 1};
        int rc = make_tensor_map_bf16(&tmBs, B, 3, dims, strides, box, true);
        if (rc) return rc;
    }
    // two row tiles per pass for grouped GEMMs whose groups hold roughly 100-256 rows (S ~ 400-1000 prompt tokens over
    // 8 experts): the second row tile of a group reuses the weight stage instead of streaming it again
    const char* mt_env = getenv("VITA_B200_GEMM_MT");   // 2 = use the two-row-tile variant for the grouped GEMMs
    const int rows_per_group = occ_rows / (args.num_groups > 0 ? args.num_groups : 1);
    // Measured (profiles/r02_gemm_shapes.txt, S = 506): 395 vs 338 us for gate|up, 204 vs 184 us for down -- the variant
    // has one ring stage less (3 x 64 KB) and, with the whole TMEM holding its two accumulators, no epilogue overlap, which
    // costs more HBM idle time than the second pass over an L2-resident weight tile it saves.  Opt-in only.
    (void)rows_per_group;
    const bool two_row_tiles = mt_env && atoi(mt_env) == 2 && block_n == 256 && args.num_groups > 1;
    if (two_row_tiles) {
        if (silu) return launch_gemm<256, true, 3, 2>(tmA, tmB, tmBs, a2, max_tiles, stream);
        return launch_gemm<256, false, 3, 2>(tmA, tmB, tmBs, a2, max_tiles, stream);
    }
    if (silu && block_n == 256) return launch_gemm<256, true, 4>(tmA, tmB, tmBs, a2, max_tiles, stream);
    if (silu) return launch_gemm<128, true, 6>(tmA, tmB, tmBs, a2, max_tiles, stream);
    if (block_n == 256) return launch_gemm<256, false, 4>(tmA, tmB, tmBs, a2, max_tiles, stream);
    return launch_gemm<128, false, 6>(tmA, tmB, tmBs, a2, max_tiles, stream);
}

}  // namespace vita

using namespace vita;

extern "C" int vita_gemm_bf16(const void* A, int64_t lda, const void* B, void* C, int64_t ldc, int64_t M, int64_t N,
                              int64_t K, const void* bias, int act, const void* colscale, const void* residual,
                              int64_t ldr, void* stream) {
    GemmArgs a{};
    a.M = static_cast<int>(M);
    a.N = static_cast<int>(N);
    a.K = static_cast<int>(K);
    a.num_groups = 1;
    a.group_offsets = nullptr;
    a.C = static_cast<__nv_bfloat16*>(C);
    a.ldc = ldc;
    a.bias = static_cast<const __nv_bfloat16*>(bias);
    a.colscale = static_cast<const __nv_bfloat16*>(colscale);
    a.rowscale = nullptr;
    a.residual = static_cast<const __nv_bfloat16*>(residual);
    a.ldr = ldr;
    a.act = act;
    VITA_REQUIRE(act == VITA_ACT_NONE || act == VITA_ACT_GELU || act == VITA_ACT_RELU, "unknown activation");
    return gemm_dispatch(A, lda, a.M, B, a.N, a, false, static_cast<cudaStream_t>(stream));
}

// qkv projection with the RoPE + paged-KV-append epilogue (see GemmArgs): qkv_out [M, (n_q + 2 n_kv) * 128] receives the
// rotated q and k heads and the v heads (what the prefill attention reads), k_cache / v_cache the rows of the slots.
extern "C" int vita_gemm_qkv_rope(const void* X, int64_t ldx, const void* W_qkv, void* qkv_out, int64_t M, int64_t K,
                                  int64_t n_q_heads, int64_t n_kv_heads, int64_t head_dim, const int32_t* positions,
                                  const int32_t* slot_mapping, const float* cos_sin, void* k_cache, void* v_cache,
                                  void* stream) {
    VITA_REQUIRE(head_dim == 128, "head_dim must be 128");
    VITA_REQUIRE(positions != nullptr && cos_sin != nullptr, "positions and cos_sin are required");
    VITA_REQUIRE(slot_mapping == nullptr || (k_cache != nullptr && v_cache != nullptr), "slot_mapping needs both caches");
    VITA_REQUIRE(aligned16(cos_sin) && (!k_cache || aligned16(k_cache)) && (!v_cache || aligned16(v_cache)), "alignment");
    GemmArgs a{};
    a.M = static_cast<int>(M);
    a.N = static_cast<int>((n_q_heads + 2 * n_kv_heads) * 128);
    a.K = static_cast<int>(K);
    a.num_groups = 1;
    a.C = static_cast<__nv_bfloat16*>(qkv_out);
    a.ldc = a.N;
    a.act = VITA_ACT_NONE;
    a.rope_cos_sin = cos_sin;
    a.rope_pos = positions;
    a.rope_slot = slot_mapping;
    a.k_cache = static_cast<__nv_bfloat16*>(k_cache);
    a.v_cache = static_cast<__nv_bfloat16*>(v_cache);
    a.n_q = static_cast<int>(n_q_heads);
    a.n_kv = static_cast<int>(n_kv_heads);
    return gemm_dispatch(X, ldx, a.M, W_qkv, a.N, a, false, static_cast<cudaStream_t>(stream));
}

extern "C" int vita_moe_gemm_gate_up_silu(const void* X_perm, const void* W_gate_up, void* Act,
                                          const int32_t* expert_offsets, int64_t rows, int64_t num_experts,
                                          int64_t H, int64_t I, void* stream) {
    GemmArgs a{};
    a.M = static_cast<int>(rows);
    a.N = static_cast<int>(I);
    a.K = static_cast<int>(H);
    a.num_groups = static_cast<int>(num_experts);
    a.group_offsets = expert_offsets;
    a.C = static_cast<__nv_bfloat16*>(Act);
    a.ldc = I;
    a.act = VITA_ACT_NONE;
    VITA_REQUIRE(expert_offsets != nullptr, "expert_offsets required");
    VITA_REQUIRE(I % 8 == 0, "I must be a multiple of 8");
    return gemm_dispatch(X_perm, H, a.M, W_gate_up, static_cast<int>(2 * I), a, true,
                         static_cast<cudaStream_t>(stream));
}

// The two grouped GEMMs over the "slot" layout of vita_moe_route_scatter: expert e owns rows [e * capacity,
// e * capacity + expert_counts[e]) of X_slots / Act_slots / Y_slots (counts on the device, no host sync).
extern "C" int vita_moe_gemm_gate_up_silu_slots(const void* X_slots, const void* W_gate_up, void* Act_slots,
                                                const int32_t* expert_counts, int64_t capacity, int64_t rows_hint,
                                                int64_t num_experts, int64_t H, int64_t I, void* stream) {
    VITA_REQUIRE(expert_counts != nullptr && capacity > 0, "expert_counts and capacity required");
    VITA_REQUIRE(I % 8 == 0 && num_experts * capacity < (1ll << 31), "I % 8 == 0 and E * capacity < 2^31");
    GemmArgs a{};
    a.M = static_cast<int>(num_experts * capacity);
    a.N = static_cast<int>(I);
    a.K = static_cast<int>(H);
    a.num_groups = static_cast<int>(num_experts);
    a.group_counts = expert_counts;
    a.group_stride = static_cast<int>(capacity);
    a.est_rows = static_cast<int>(rows_hint);
    a.C = static_cast<__nv_bfloat16*>(Act_slots);
    a.ldc = I;
    a.act = VITA_ACT_NONE;
    return gemm_dispatch(X_slots, H, a.M, W_gate_up, static_cast<int>(2 * I), a, true, static_cast<cudaStream_t>(stream));
}

extern "C" int vita_moe_gemm_down_slots(const void* Act_slots, const void* W_down, void* Y_slots,
                                        const int32_t* expert_counts, const float* row_weight, int64_t capacity,
                                        int64_t rows_hint, int64_t num_experts, int64_t H, int64_t I, void* stream) {
    VITA_REQUIRE(expert_counts != nullptr && capacity > 0, "expert_counts and capacity required");
    VITA_REQUIRE(num_experts * capacity < (1ll << 31), "E * capacity < 2^31");
    GemmArgs a{};
    a.M = static_cast<int>(num_experts * capacity);
    a.N = static_cast<int>(H);
    a.K = static_cast<int>(I);
    a.num_groups = static_cast<int>(num_experts);
    a.group_counts = expert_counts;
    a.group_stride = static_cast<int>(capacity);
    a.est_rows = static_cast<int>(rows_hint);
    a.C = static_cast<__nv_bfloat16*>(Y_slots);
    a.ldc = H;
    a.rowscale = row_weight;
    a.act = VITA_ACT_NONE;
    return gemm_dispatch(Act_slots, I, a.M, W_down, static_cast<int>(H), a, false, static_cast<cudaStream_t>(stream));
}

// Expert-parallel down projection: the GEMM epilogue stores every output row straight into the symmetric-memory
// receive buffer of the rank that owns the token (P2P stores over NVLink), so compute and the "combine" transfer are
// one kernel; the owner later sums its two slots per token (vita_ep_reduce_norm_gather).
extern "C" int vita_moe_gemm_down_ep(const void* Act, const void* W_down, const int32_t* expert_offsets,
                                     const float* row_weight, const int32_t* row_assign, void* const* peer_out,
                                     int64_t rows, int64_t num_local_experts, int64_t H, int64_t I, int64_t chunk,
                                     void* stream) {
    GemmArgs a{};
    a.M = static_cast<int>(rows);
    a.N = static_cast<int>(H);
    a.K = static_cast<int>(I);
    a.num_groups = static_cast<int>(num_local_experts);
    a.group_offsets = expert_offsets;
    a.C = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(Act));   // unused when peer_out is set; keeps checks happy
    a.ldc = H;
    a.rowscale = row_weight;
    a.act = VITA_ACT_NONE;
    a.row_assign = row_assign;
    a.peer_out = reinterpret_cast<__nv_bfloat16* const*>(peer_out);
    a.chunk = static_cast<int>(chunk);
    VITA_REQUIRE(expert_offsets && row_assign && peer_out && chunk > 0, "offsets, row_assign, peer_out and chunk required");
    return gemm_dispatch(Act, I, a.M, W_down, static_cast<int>(H), a, false, static_cast<cudaStream_t>(stream));
}

extern "C" int vita_moe_gemm_down(const void* Act, const void* W_down, void* Y_perm, const int32_t* expert_offsets,
                                  const float* row_weight, int64_t rows, int64_t num_experts, int64_t H, int64_t I,
                                  void* stream) {
    GemmArgs a{};
    a.M = static_cast<int>(rows);
    a.N = static_cast<int>(H);
    a.K = static_cast<int>(I);
    a.num_groups = static_cast<int>(num_experts);
    a.group_offsets = expert_offsets;
    a.C = static_cast<__nv_bfloat16*>(Y_perm);
    a.ldc = H;
    a.rowscale = row_weight;
    a.act = VITA_ACT_NONE;
    VITA_REQUIRE(expert_offsets != nullptr, "expert_offsets required");
    return gemm_dispatch(Act, I, a.M, W_down, static_cast<int>(H), a, false, static_cast<cudaStream_t>(stream));
}
