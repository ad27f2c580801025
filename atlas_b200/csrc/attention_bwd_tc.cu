// EXPERIMENTAL (selected only with ATLAS_B200_ATTN_BWD_TC=1; the default backward is the warp-MMA path of
// attention_bwd.cu): the dQ half of the attention backward on tcgen05, structured like the forward kernel of attention.cu
// (DESIGN.md §8 item 1).  Status at the end of round 1 (1 x B200): CORRECT - all 69 backward / training parity tests pass
// with it enabled, dQ / dbias within one 16-bit ulp of the warp-MMA kernels (tools/try_tc_bwd.py) - but not yet faster:
// 0.57 ms vs 0.52 ms for `attn_bwd_dq2_kernel` at 80 x 12 x 384 x 384, i.e. ~17 k cycles per 128 x 128 chunk against
// ~0.8 k cycles of MMA work: the chunk loop is a serial chain of four hand-offs (dP commit -> TMEM loads -> dS store ->
// dQ MMA) with only S prefetched.  Hint from the dK / dV twin (attention_bwd_tc_dkv.cu: same serial structure, TWO TS
// products and four tcgen05.st per chunk, yet ~0.23 ms): what this kernel has and the twin has not is (a) the per-tile
// D_i = dO . O dot products read from global memory by every dS thread, (b) the dbias staging + shared reductions, (c) the
// dsum stores - look there first (D as a per-item table filled by the auxiliary warp, dbias in a separate pass over a
// stored dS).  Then: double-buffer dP and dS so the dS math of chunk c+1 overlaps the dQ MMAs of chunk c, 16 dS warps
// (4 threads per row), K / V of the next item prefetched.
//
//   CTA = persistent over (segment b, head h) items; K and V of the item resident in shared memory (TMA, K-major, 128B
//   swizzle); per 128-query tile Q and dO stream in (double buffered).  Per 128-key chunk c:
//       S_c  = Q  . K_c^T      SS MMA -> tensor memory, fp32, 128 columns (double buffered: S0 | S1)
//       dP_c = dO . V_c^T      SS MMA -> tensor memory, fp32, 128 columns
//       dS_c = P o (dP - D), P = 2^(t - lse * log2e), t = scale2 * S + bias2 + mask2      (8 warps: 2 threads per query row)
//              packed to 16 bits and written to its own 64 columns with tcgen05.st
//       dQ  += dS_c . K_c      TS MMA: A from tensor memory, B = the K rows as they lie in shared memory (MN-major)
//   TMEM columns: S0 [0,128) | S1 [128,256) | dP [256,384) | dQ [384,448) | dS [448,512).
//   dbias[h, j - i] += dS[i, j]: each warp stages its 32 x 32 piece of dS in shared memory and sums its 63 diagonals (one
//   shared fp32 atomic per element was 20 k cycles per chunk: `atomicAdd(float)` on shared memory is a CAS loop); the
//   per-item table is flushed to global memory once per item.
// Also writes D_i = sum_d dO[i, d] O[i, d] (`dsum`) for the dK / dV kernel.
// Roles (352 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator + per-item mask / bias tables,
// warps 3-10 dS + output (warp w owns TMEM lanes 32 (w % 4) .., column half (w - 3) / 4).
#include "common.cuh"
#include "host_common.h"

#include <math.h>
#include <stdlib.h>

namespace attnb_tc {

constexpr int D = 64;
constexpr int BLOCK_Q = 128;
constexpr int MAX_LK = 512;
constexpr int SPLIT = 2;
constexpr int SM_THREADS = 128 * SPLIT;
constexpr int THREADS = 96 + SM_THREADS;          // 11 warps
constexpr int AUX_THREADS = 32;
constexpr int Q_BYTES = BLOCK_Q * D * 2;          // 16 KB (Q tile; the dO tile has the same size)
constexpr int KV_BYTES = MAX_LK * D * 2;          // 64 KB each
constexpr int SMEM_BYTES = 4 * Q_BYTES + 2 * KV_BYTES + 1024;   // Q[2] | dO[2] | K | V
constexpr int TMEM_COLS = 512;
constexpr uint32_t COL_DP = 256, COL_DQ = 384, COL_DS = 448;
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
    int B, H, Lq, Lk;
    int q_col0, k_col0, v_col0;
    const uint16_t *o, *dout;
    int64_t ldo, lddo;
    uint16_t* dq;
    int64_t lddq;
    int dq_col0;
    const float* add_mask;
    const float* bias_delta;
    float* dbias;
    const float* lse;
    float* dsum;
    float scale, causal_value;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// MN-major operand tile: rows of 128 bytes (64 x 16-bit along MN), 8-row K groups 1024 bytes apart (see attention.cu)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1024 >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ void sm_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SM_THREADS) : "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool kBF16>
__device__ __forceinline__ float to_f32(uint32_t h) {
    if constexpr (kBF16) return __bfloat162float(__ushort_as_bfloat16(static_cast<unsigned short>(h & 0xFFFFu)));
    return __half2float(__ushort_as_half(static_cast<unsigned short>(h & 0xFFFFu)));
}

template <bool kBF16>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                      const Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t k_full, v_full, kv_empty, q_full[2], q_empty[2], s_full[2], dp_full, ds_ready, ds_free,
        dq_full, dq_free;
    __shared__ __align__(8) uint64_t aux_full[2], aux_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_bias[2][2 * MAX_LK];             // (bias by (j - i) + (Lq - 1) [+ causal]) * log2e
    __shared__ __align__(16) float s_mask[2][MAX_LK];   // additive key mask * log2e (-inf beyond Lk)
    __shared__ float s_dbias[2 * MAX_LK];               // per-item dbias by (j - i) + (Lq - 1)
    __shared__ uint32_t s_stage[8][32 * 16];            // per softmax warp: one 32 x 32 dS piece, packed 16-bit pairs

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t smem_base = (ab::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ab::smem_u32(smem_raw));
    uint8_t* sQ = smem_gen;                          // [2][16 KB]
    uint8_t* sdO = smem_gen + 2 * Q_BYTES;           // [2][16 KB]
    uint8_t* sK = smem_gen + 4 * Q_BYTES;
    uint8_t* sV = sK + KV_BYTES;
    const uint32_t aQ = smem_base, adO = smem_base + 2 * Q_BYTES, aK = smem_base + 4 * Q_BYTES, aV = aK + KV_BYTES;

    const int n_chunks = (p.Lk + 127) / 128;
    const int lk_pad = n_chunks * 128;
    const int n_qt = (p.Lq + BLOCK_Q - 1) / BLOCK_Q;
    const int n_items = p.B * p.H;
    const int ntab = p.Lq + p.Lk - 1;

    if (warp == 0 && lane == 0) {
        ab::tma_prefetch_desc(&tmap_q);
        ab::tma_prefetch_desc(&tmap_k);
        ab::tma_prefetch_desc(&tmap_v);
        ab::tma_prefetch_desc(&tmap_do);
    }
    if (warp == 1 && lane == 0) {
        ab::mbar_init(&k_full, 1);
        ab::mbar_init(&v_full, 1);
        ab::mbar_init(&kv_empty, 1);
        for (int i = 0; i < 2; ++i) {
            ab::mbar_init(&q_full[i], 1);
            ab::mbar_init(&q_empty[i], 1);
            ab::mbar_init(&s_full[i], 1);
            ab::mbar_init(&aux_full[i], AUX_THREADS);
            ab::mbar_init(&aux_empty[i], SM_THREADS);
        }
        ab::mbar_init(&dp_full, 1);
        ab::mbar_init(&ds_ready, SM_THREADS);
        ab::mbar_init(&ds_free, 1);
        ab::mbar_init(&dq_full, 1);
        ab::mbar_init(&dq_free, SM_THREADS);
        ab::fence_barrier_init();
    }
    if (warp == 2) ab::tmem_alloc<1>(&tmem_base_smem, TMEM_COLS);
    ab::tc_fence_before();
    __syncthreads();
    ab::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int item_it = 0, qt_it = 0;
            const uint32_t kv_bytes = static_cast<uint32_t>(n_chunks * 128 * D * 2);
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
                const int b = item / p.H, h = item % p.H;
                ab::mbar_wait(&kv_empty, (item_it & 1) ^ 1u, 61);
                ab::mbar_arrive_expect_tx(&k_full, kv_bytes);
                for (int c = 0; c < n_chunks; ++c)
                    ab::tma_load_2d(&tmap_k, &k_full, sK + c * (128 * D * 2), p.k_col0 + h * D, b * p.Lk + c * 128,
                                    ab::kEvictNormal);
                ab::mbar_arrive_expect_tx(&v_full, kv_bytes);
                for (int c = 0; c < n_chunks; ++c)
                    ab::tma_load_2d(&tmap_v, &v_full, sV + c * (128 * D * 2), p.v_col0 + h * D, b * p.Lk + c * 128,
                                    ab::kEvictNormal);
                for (int qt = 0; qt < n_qt; ++qt, ++qt_it) {
                    const int qb = qt_it & 1;
                    ab::mbar_wait(&q_empty[qb], ((qt_it >> 1) & 1) ^ 1u, 62);
                    ab::mbar_arrive_expect_tx(&q_full[qb], 2 * Q_BYTES);
                    ab::tma_load_2d(&tmap_q, &q_full[qb], sQ + qb * Q_BYTES, p.q_col0 + h * D, b * p.Lq + qt * BLOCK_Q,
                                    ab::kEvictFirst);
                    ab::tma_load_2d(&tmap_do, &q_full[qb], sdO + qb * Q_BYTES, h * D, b * p.Lq + qt * BLOCK_Q,
                                    ab::kEvictFirst);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = ab::umma_idesc_f16(BLOCK_Q, 128, kBF16);
            constexpr uint32_t idesc_dq = ab::umma_idesc_f16(BLOCK_Q, D, kBF16) | (1u << 16);   // B = K rows, MN-major
            int item_it = 0, qt_it = 0, ch = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
                ab::mbar_wait(&k_full, item_it & 1, 63);
                ab::mbar_wait(&v_full, item_it & 1, 64);
                for (int qt = 0; qt < n_qt; ++qt, ++qt_it) {
                    const int qb = qt_it & 1;
                    ab::mbar_wait(&q_full[qb], (qt_it >> 1) & 1, 65);
                    ab::tc_fence_after();
                    const uint64_t qdesc = ab::umma_desc_k_sw128(aQ + qb * Q_BYTES);
                    const uint64_t dodesc = ab::umma_desc_k_sw128(adO + qb * Q_BYTES);
                    auto issue_s = [&](int c, int chi) {
                        const uint64_t kdesc = ab::umma_desc_k_sw128(aK + c * (128 * D * 2));
#pragma unroll
                        for (int k = 0; k < D / 16; ++k)
                            ab::umma_ss<1>(tmem_base + (chi & 1) * 128, qdesc + ((k * 32) >> 4), kdesc + ((k * 32) >> 4),
                                           idesc_s, k != 0 ? 1u : 0u);
                        ab::umma_commit(&s_full[chi & 1]);
                    };
                    auto issue_dp = [&](int c) {
                        const uint64_t vdesc = ab::umma_desc_k_sw128(aV + c * (128 * D * 2));
#pragma unroll
                        for (int k = 0; k < D / 16; ++k)
                            ab::umma_ss<1>(tmem_base + COL_DP, dodesc + ((k * 32) >> 4), vdesc + ((k * 32) >> 4), idesc_s,
                                           k != 0 ? 1u : 0u);
                        ab::umma_commit(&dp_full);
                    };
                    issue_s(0, ch);
                    issue_dp(0);
                    for (int c = 0; c < n_chunks; ++c, ++ch) {
                        if (c + 1 < n_chunks) issue_s(c + 1, ch + 1);     // the other S buffer: S_{c-1} was consumed
                        ab::mbar_wait(&ds_ready, ch & 1, 66);             // dS_c is in tensor memory; S_c, dP_c were read
                        if (c == 0) ab::mbar_wait(&dq_free, (qt_it & 1) ^ 1u, 67);   // previous tile's dQ was read out
                        ab::tc_fence_after();
                        const uint64_t kmn = umma_desc_mn_sw128(aK + c * (128 * D * 2));
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            ab::umma_ts<1>(tmem_base + COL_DQ, tmem_base + COL_DS + k * 8,
                                           kmn + static_cast<uint64_t>((k * 2048) >> 4), idesc_dq, (c | k) != 0 ? 1u : 0u);
                        ab::umma_commit(&ds_free);
                        if (c == n_chunks - 1) ab::umma_commit(&dq_full);
                        if (c + 1 < n_chunks) issue_dp(c + 1);
                    }
                    ab::umma_commit(&q_empty[qb]);      // every MMA reading this tile's Q / dO has been issued
                }
                ab::umma_commit(&kv_empty);             // ... and this item's K / V
            }
        }
    } else if (warp == 2) {
        // ===================== per-item mask / bias tables (one item ahead) =====================
        const int tid = static_cast<int>(lane);
        const bool has_bias = (p.bias_delta != nullptr) || (p.causal_value != 0.f);
        int item_it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
            const int b = item / p.H, h = item % p.H;
            const int buf = item_it & 1;
            ab::mbar_wait(&aux_empty[buf], ((item_it >> 1) & 1) ^ 1u, 68);
            for (int j = tid; j < lk_pad; j += AUX_THREADS)
                s_mask[buf][j] = (j < p.Lk) ? (p.add_mask ? p.add_mask[static_cast<size_t>(b) * p.Lk + j] * LOG2E : 0.f)
                                            : -INFINITY;
            if (has_bias)
                for (int d = tid; d < 2 * MAX_LK; d += AUX_THREADS) {
                    float v = 0.f;
                    if (d < ntab) {
                        v = p.bias_delta ? p.bias_delta[static_cast<size_t>(h) * ntab + d] : 0.f;
                        if (p.causal_value != 0.f && d > p.Lq - 1) v += p.causal_value;
                    }
                    s_bias[buf][d] = v * LOG2E;
                }
            ab::mbar_arrive(&aux_full[buf]);
        }
    } else {
        // ===================== dS + output: two threads per query row =====================
        const uint32_t lg = warp & 3u;                       // TMEM lane group of this warp
        const uint32_t part = (warp - 3u) >> 2;              // which 64 of the chunk's 128 key columns
        const int row = static_cast<int>(lg * 32 + lane);
        const int sm_tid = static_cast<int>(threadIdx.x) - 96;
        const uint32_t lane_addr = tmem_base + ((lg * 32u) << 16);
        const float scale2 = p.scale * LOG2E;
        const bool has_bias = (p.bias_delta != nullptr) || (p.causal_value != 0.f);
        const bool want_dbias = p.dbias != nullptr;
        int item_it = 0, qt_it = 0, ch = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
            const int b = item / p.H, h = item % p.H;
            const int buf = item_it & 1;
            if (want_dbias) {
                for (int x = sm_tid; x < ntab; x += SM_THREADS) s_dbias[x] = 0.f;
                sm_bar();
            }
            ab::mbar_wait(&aux_full[buf], (item_it >> 1) & 1, 69);
            const float* mask2 = s_mask[buf];
            const float* bias2 = s_bias[buf];
            for (int qt = 0; qt < n_qt; ++qt, ++qt_it) {
                const int i = qt * BLOCK_Q + row;
                const bool live = i < p.Lq;
                const int boff = min(p.Lq - 1 - i, 2 * MAX_LK - 1 - lk_pad);
                const int boffc = max(boff, 0);
                // D_i and lse of this row (padding rows: lse = +inf -> every probability is 0)
                float drow = 0.f, lse2 = INFINITY;
                if (live) {
                    const int64_t grow = static_cast<int64_t>(b) * p.Lq + i;
                    const uint4* po = reinterpret_cast<const uint4*>(p.o + grow * p.ldo + h * D);
                    const uint4* pd = reinterpret_cast<const uint4*>(p.dout + grow * p.lddo + h * D);
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        const uint4 a = __ldg(po + c8), d = __ldg(pd + c8);
                        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            drow = fmaf(to_f32<kBF16>(aw[e]), to_f32<kBF16>(dw[e]), drow);
                            drow = fmaf(to_f32<kBF16>(aw[e] >> 16), to_f32<kBF16>(dw[e] >> 16), drow);
                        }
                    }
                    const int64_t si = (static_cast<int64_t>(b) * p.H + h) * p.Lq + i;
                    lse2 = __ldg(p.lse + si) * LOG2E;
                    if (part == 0) p.dsum[si] = drow;
                }
                for (int c = 0; c < n_chunks; ++c, ++ch) {
                    ab::mbar_wait(&s_full[ch & 1], (ch >> 1) & 1, 70);
                    ab::mbar_wait(&dp_full, ch & 1, 71);
                    ab::tc_fence_after();
                    uint32_t pk[2][16];
#pragma unroll
                    for (int piece = 0; piece < 2; ++piece) {
                        const int col0 = static_cast<int>(part) * 64 + piece * 32;      // within the chunk
                        const int j0 = c * 128 + col0;                                  // key position of column 0
                        uint32_t rs[32], rd[32];
                        ab::tmem_ld32(lane_addr + (ch & 1) * 128 + col0, rs);
                        ab::tmem_ld32(lane_addr + COL_DP + col0, rd);
                        ab::tmem_ld_wait();
                        const float4* m4 = reinterpret_cast<const float4*>(mask2 + j0);
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {
                            const float4 m = m4[q4];
                            const float add[4] = {m.x, m.y, m.z, m.w};
                            float ds[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int jj = 4 * q4 + e;
                                float a = add[e];
                                if (has_bias) a += bias2[j0 + jj + boffc];
                                const float t = fmaf(__uint_as_float(rs[jj]), scale2, a);
                                const float pr = ex2_approx(t - lse2);
                                ds[e] = pr * (__uint_as_float(rd[jj]) - drow);      // 0 for padding rows / keys (pr = 0)
                            }
                            pk[piece][2 * q4] = ab::pack2_rn<kBF16>(ds[0], ds[1]);
                            pk[piece][2 * q4 + 1] = ab::pack2_rn<kBF16>(ds[2], ds[3]);
                        }
                        if (want_dbias) {
                            // dbias[h, j - i] += dS[i, j] over this warp's 32 x 32 piece (rows = lanes): the packed piece goes
                            // through a per-warp shared tile rotated by the row (word(l, w) = 16 l + ((w + l) & 15): stores and
                            // diagonal reads are conflict free), lane d sums diagonals d and d - 32, then two shared
                            // reductions per lane instead of 32
                            uint32_t* st = s_stage[warp - 3u];
#pragma unroll
                            for (int w = 0; w < 16; ++w) st[lane * 16 + ((w + lane) & 15)] = pk[piece][w];
                            __syncwarp();
                            const int i0 = qt * BLOCK_Q + static_cast<int>(lg) * 32;
                            const int base = j0 - i0 + p.Lq - 1;        // table index of (row i0, column j0)
#pragma unroll
                            for (int sgn = 0; sgn < 2; ++sgn) {
                                const int dl = static_cast<int>(lane) - 32 * sgn;       // diagonal c - r
                                const int r_lo = dl < 0 ? -dl : 0, r_hi = dl > 0 ? 31 - dl : 31;
                                float sum = 0.f;
                                for (int r = r_lo; r <= r_hi; ++r) {
                                    const int c = r + dl;
                                    const uint32_t wv = st[r * 16 + (((c >> 1) + r) & 15)];
                                    sum += to_f32<kBF16>((c & 1) ? (wv >> 16) : wv);
                                }
                                const int idx = base + dl;
                                if (sum != 0.f && idx >= 0 && idx < ntab) atomicAdd(&s_dbias[idx], sum);
                            }
                            __syncwarp();
                        }
                    }
                    ab::mbar_wait(&ds_free, (ch & 1) ^ 1u, 72);      // the dQ MMAs of the previous chunk have read dS
                    ab::tc_fence_after();
                    tmem_st16(lane_addr + COL_DS + part * 32, pk[0]);
                    tmem_st16(lane_addr + COL_DS + part * 32 + 16, pk[1]);
                    ab::tmem_st_wait();
                    ab::tc_fence_before();
                    ab::mbar_arrive(&ds_ready);
                }
                // ---- this tile's dQ: 32 of the 64 columns per thread ----
                ab::mbar_wait(&dq_full, qt_it & 1, 73);
                ab::tc_fence_after();
                {
                    uint32_t r[32];
                    ab::tmem_ld32(lane_addr + COL_DQ + part * 32, r);
                    ab::tmem_ld_wait();
                    if (live) {
                        uint4* dst = reinterpret_cast<uint4*>(p.dq + (static_cast<int64_t>(b) * p.Lq + i) * p.lddq + p.dq_col0 +
                                                              h * D + part * 32);
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4)
                            dst[v4] = make_uint4(
                                ab::pack2_rn<kBF16>(__uint_as_float(r[8 * v4]) * p.scale, __uint_as_float(r[8 * v4 + 1]) * p.scale),
                                ab::pack2_rn<kBF16>(__uint_as_float(r[8 * v4 + 2]) * p.scale, __uint_as_float(r[8 * v4 + 3]) * p.scale),
                                ab::pack2_rn<kBF16>(__uint_as_float(r[8 * v4 + 4]) * p.scale, __uint_as_float(r[8 * v4 + 5]) * p.scale),
                                ab::pack2_rn<kBF16>(__uint_as_float(r[8 * v4 + 6]) * p.scale, __uint_as_float(r[8 * v4 + 7]) * p.scale));
                    }
                }
                ab::tc_fence_before();
                ab::mbar_arrive(&dq_free);
            }
            ab::mbar_arrive(&aux_empty[buf]);       // last read of this item's tables
            if (want_dbias) {
                sm_bar();
                for (int x = sm_tid; x < ntab; x += SM_THREADS) {
                    const float v = s_dbias[x];
                    if (v != 0.f) atomicAdd(p.dbias + static_cast<size_t>(h) * ntab + x, v);
                }
                sm_bar();
            }
        }
    }

    ab::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ab::tc_fence_after();
        ab::tmem_dealloc<1>(tmem_base, TMEM_COLS);
    }
}

}  // namespace attnb_tc

// Launched by atlas_b200_attention_bwd (attention_bwd.cu) when ATLAS_B200_ATTN_BWD_TC=1 and the shape qualifies
// (Lk <= 512, Lq + Lk - 1 <= 1024, forward lse available, keys not split).  Returns ATLAS_B200_EUNSUPPORTED otherwise.
int atlas_b200_attn_bwd_dq_tc(const void* q, int64_t ldq, int32_t q_col0, const void* k, int64_t ldk, int32_t k_col0,
                              const void* v, int64_t ldv, int32_t v_col0, const void* out, int64_t ldo, const void* dout,
                              int64_t lddo, void* dq, int64_t lddq, int32_t dq_col0, const float* add_mask,
                              const float* bias_delta, float* dbias_delta, const float* lse, float* dsum, int32_t B,
                              int32_t H, int32_t Lq, int32_t Lk, float scale, float causal_value, int32_t is_bf16,
                              cudaStream_t s) {
    using namespace attnb_tc;
    if (Lk > MAX_LK || Lq + Lk - 1 > 2 * MAX_LK) return ATLAS_B200_EUNSUPPORTED;
    CUtensorMap tq, tk, tv, tdo;
    int rc = abh::make_tmap_2d_16bit(&tq, q, static_cast<uint64_t>(B) * Lq, static_cast<uint64_t>(q_col0 + H * D),
                                     static_cast<uint64_t>(ldq), BLOCK_Q, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tk, k, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(k_col0 + H * D),
                                 static_cast<uint64_t>(ldk), 128, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tv, v, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(v_col0 + H * D),
                                 static_cast<uint64_t>(ldv), 128, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tdo, dout, static_cast<uint64_t>(B) * Lq, static_cast<uint64_t>(H * D),
                                 static_cast<uint64_t>(lddo), BLOCK_Q, D, is_bf16 != 0);
    if (rc) return rc;
    Params p;
    p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.o = static_cast<const uint16_t*>(out);
    p.dout = static_cast<const uint16_t*>(dout);
    p.ldo = ldo; p.lddo = lddo;
    p.dq = static_cast<uint16_t*>(dq);
    p.lddq = lddq;
    p.dq_col0 = dq_col0;
    p.add_mask = add_mask;
    p.bias_delta = bias_delta;
    p.dbias = dbias_delta;
    p.lse = lse;
    p.dsum = dsum;
    p.scale = scale;
    p.causal_value = causal_value;
    const int items = B * H;
    const int grid = items < abh::num_sms() ? items : abh::num_sms();
    static bool attr_set[2] = {false, false};
    if (is_bf16) {
        if (!attr_set[0]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set[0] = true;
        }
        attn_bwd_dq_tc_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    } else {
        if (!attr_set[1]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set[1] = true;
        }
        attn_bwd_dq_tc_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}
