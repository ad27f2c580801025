// EXPERIMENTAL (selected only with ATLAS_B200_ATTN_BWD_TC=2; the default backward is the warp-MMA path of attention_bwd.cu):
// the dK / dV half of the attention backward on tcgen05, the transposed twin of attn_bwd_dq_tc_kernel
// (attention_bwd_tc.cu).  Status at the end of round 1 (1 x B200, the last seconds of the round's GPU budget): numerically
// right on the four shapes of tools/try_tc_bwd.py (bf16 / fp16, L = 64 / 200 / 384, with mask and relative-position bias):
// dK and dV within one 16-bit ulp of the warp-MMA kernels, and ~0.23 ms against 0.33 ms for `attn_bwd_dkv2_kernel` at
// 80 x 12 x 384 x 384 (total backward with both tcgen05 kernels 0.80 ms vs 0.85 ms).  NOT yet run under the full parity
// suite (causal, L = 7 / 100 / 130 cases) - round 2 starts with
// `ATLAS_B200_ATTN_BWD_TC=2 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py -m gpu`.
//
//   CTA = persistent over (segment b, head h) items; Q and dO of the item resident in shared memory (128-row chunks, TMA,
//   K-major, 128B swizzle); per 128-key tile K_t and V_t stream in (double buffered).  Per 128-query chunk c:
//       S^T_c  = K_t . Q_c^T       SS MMA -> tensor memory [keys on lanes, queries on columns], fp32, 128 columns
//       dP^T_c = V_t . dO_c^T      SS MMA -> 128 columns
//       P^T = 2^(t - lse2[i]), dS^T = P^T o (dP^T - D[i])            (8 warps: 2 threads per key row), both packed to 16 bits
//       dV += P^T  . dO_c          TS MMA: A from tensor memory, B = dO rows as they lie in shared memory (MN-major)
//       dK += dS^T . Q_c           TS MMA: A from tensor memory, B = Q rows (MN-major)
//   TMEM columns: S^T [0,128) | dP^T [128,256) | P^T [256,320) | dS^T [320,384) | dV [384,448) | dK [448,512).
//   lse and D (written by the dQ kernel, which must run first) are per-item tables in shared memory, indexed by the
//   query column; the additive key mask is a per-thread scalar.
// Roles (352 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator + per-item tables, warps 3-10 the
// dS math + output (warp w owns TMEM lanes 32 (w % 4) .., query-column half (w - 3) / 4).
#include "common.cuh"
#include "host_common.h"

#include <math.h>
#include <stdlib.h>

namespace attnb_tc_dkv {

constexpr int D = 64;
constexpr int BLOCK = 128;                        // keys per tile = queries per chunk
constexpr int MAX_L = 512;
constexpr int SPLIT = 2;
constexpr int SM_THREADS = 128 * SPLIT;
constexpr int THREADS = 96 + SM_THREADS;
constexpr int AUX_THREADS = 32;
constexpr int TILE_BYTES = BLOCK * D * 2;         // 16 KB
constexpr int QDO_BYTES = MAX_L * D * 2;          // 64 KB each: Q and dO of the whole segment
constexpr int SMEM_BYTES = 2 * QDO_BYTES + 4 * TILE_BYTES + 1024;   // Q | dO | K[2] | V[2]
constexpr int TMEM_COLS = 512;
constexpr uint32_t COL_DPT = 128, COL_PT = 256, COL_DST = 320, COL_DV = 384, COL_DK = 448;
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
    int B, H, Lq, Lk;
    int q_col0, k_col0, v_col0;
    uint16_t *dk, *dv;
    int64_t lddk, lddv;
    int dk_col0, dv_col0;
    const float* add_mask;
    const float* bias_delta;
    const float* lse;
    const float* dsum;
    const uint8_t* blk_live;     // [B, ceil(Lk / 64)] or nullptr: key blocks whose keys are all masked out (probability exactly 0
                                 // in fp32 -> dK = dV = 0 exactly): a 128-key tile of two dead blocks is skipped by every role
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

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1024 >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool kBF16>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                       const Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t qdo_full, qdo_empty, kv_full[2], kv_empty[2], sp_full, pd_ready, pd_free, dkv_full, dkv_free;
    __shared__ __align__(8) uint64_t aux_full[2], aux_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_bias[2][2 * MAX_L];               // (bias by (j - i) + (Lq - 1) [+ causal]) * log2e
    __shared__ float s_mask[2][MAX_L];                   // additive key mask * log2e (-inf beyond Lk)
    __shared__ __align__(16) float s_lse[2][MAX_L];      // lse * log2e per query (+inf beyond Lq: probability 0)
    __shared__ __align__(16) float s_d[2][MAX_L];        // D_i = sum_d dO O per query (0 beyond Lq)

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t smem_base = (ab::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ab::smem_u32(smem_raw));
    uint8_t* sQ = smem_gen;
    uint8_t* sdO = smem_gen + QDO_BYTES;
    uint8_t* sK = smem_gen + 2 * QDO_BYTES;              // [2][16 KB]
    uint8_t* sV = sK + 2 * TILE_BYTES;                   // [2][16 KB]
    const uint32_t aQ = smem_base, adO = smem_base + QDO_BYTES, aK = smem_base + 2 * QDO_BYTES, aV = aK + 2 * TILE_BYTES;

    const int n_kt = (p.Lk + BLOCK - 1) / BLOCK;         // key tiles
    const int n_qc = (p.Lq + BLOCK - 1) / BLOCK;         // query chunks
    const int lq_pad = n_qc * BLOCK, lk_pad = n_kt * BLOCK;
    const int n_items = p.B * p.H;
    const int ntab = p.Lq + p.Lk - 1;
    const int nb64 = (p.Lk + 63) / 64;
    // identical in every role (uniform global reads): the barrier phase counters only advance for live tiles
    auto tile_dead = [&](int b, int kt) -> bool {
        if (p.blk_live == nullptr) return false;
        const uint8_t* f = p.blk_live + static_cast<size_t>(b) * nb64 + 2 * kt;
        return f[0] == 0 && (2 * kt + 1 >= nb64 || f[1] == 0);
    };

    if (warp == 0 && lane == 0) {
        ab::tma_prefetch_desc(&tmap_q);
        ab::tma_prefetch_desc(&tmap_k);
        ab::tma_prefetch_desc(&tmap_v);
        ab::tma_prefetch_desc(&tmap_do);
    }
    if (warp == 1 && lane == 0) {
        ab::mbar_init(&qdo_full, 1);
        ab::mbar_init(&qdo_empty, 1);
        for (int i = 0; i < 2; ++i) {
            ab::mbar_init(&kv_full[i], 1);
            ab::mbar_init(&kv_empty[i], 1);
            ab::mbar_init(&aux_full[i], AUX_THREADS);
            ab::mbar_init(&aux_empty[i], SM_THREADS);
        }
        ab::mbar_init(&sp_full, 1);
        ab::mbar_init(&pd_ready, SM_THREADS);
        ab::mbar_init(&pd_free, 1);
        ab::mbar_init(&dkv_full, 1);
        ab::mbar_init(&dkv_free, SM_THREADS);
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
            int item_it = 0, kt_it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
                const int b = item / p.H, h = item % p.H;
                ab::mbar_wait(&qdo_empty, (item_it & 1) ^ 1u, 81);
                ab::mbar_arrive_expect_tx(&qdo_full, static_cast<uint32_t>(2 * n_qc * TILE_BYTES));
                for (int c = 0; c < n_qc; ++c) {
                    ab::tma_load_2d(&tmap_q, &qdo_full, sQ + c * TILE_BYTES, p.q_col0 + h * D, b * p.Lq + c * BLOCK,
                                    ab::kEvictNormal);
                    ab::tma_load_2d(&tmap_do, &qdo_full, sdO + c * TILE_BYTES, h * D, b * p.Lq + c * BLOCK, ab::kEvictNormal);
                }
                for (int kt = 0; kt < n_kt; ++kt) {
                    if (tile_dead(b, kt)) continue;
                    const int kb = kt_it & 1;
                    ab::mbar_wait(&kv_empty[kb], ((kt_it >> 1) & 1) ^ 1u, 82);
                    ab::mbar_arrive_expect_tx(&kv_full[kb], 2 * TILE_BYTES);
                    ab::tma_load_2d(&tmap_k, &kv_full[kb], sK + kb * TILE_BYTES, p.k_col0 + h * D, b * p.Lk + kt * BLOCK,
                                    ab::kEvictFirst);
                    ab::tma_load_2d(&tmap_v, &kv_full[kb], sV + kb * TILE_BYTES, p.v_col0 + h * D, b * p.Lk + kt * BLOCK,
                                    ab::kEvictFirst);
                    ++kt_it;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = ab::umma_idesc_f16(BLOCK, BLOCK, kBF16);
            constexpr uint32_t idesc_o = ab::umma_idesc_f16(BLOCK, D, kBF16) | (1u << 16);   // B = dO / Q rows, MN-major
            int item_it = 0, kt_it = 0, ch = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
                ab::mbar_wait(&qdo_full, item_it & 1, 83);
                const int b = item / p.H;
                for (int kt = 0; kt < n_kt; ++kt) {
                    if (tile_dead(b, kt)) continue;
                    const int kb = kt_it & 1;
                    ab::mbar_wait(&kv_full[kb], (kt_it >> 1) & 1, 84);
                    ab::tc_fence_after();
                    const uint64_t kdesc = ab::umma_desc_k_sw128(aK + kb * TILE_BYTES);
                    const uint64_t vdesc = ab::umma_desc_k_sw128(aV + kb * TILE_BYTES);
                    auto issue_sp = [&](int c) {
                        const uint64_t qdesc = ab::umma_desc_k_sw128(aQ + c * TILE_BYTES);
                        const uint64_t dodesc = ab::umma_desc_k_sw128(adO + c * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < D / 16; ++k)
                            ab::umma_ss<1>(tmem_base, kdesc + ((k * 32) >> 4), qdesc + ((k * 32) >> 4), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < D / 16; ++k)
                            ab::umma_ss<1>(tmem_base + COL_DPT, vdesc + ((k * 32) >> 4), dodesc + ((k * 32) >> 4), idesc_s,
                                           k != 0 ? 1u : 0u);
                        ab::umma_commit(&sp_full);
                    };
                    issue_sp(0);
                    for (int c = 0; c < n_qc; ++c, ++ch) {
                        ab::mbar_wait(&pd_ready, ch & 1, 85);       // P^T, dS^T are in tensor memory; S^T, dP^T were read
                        if (c == 0) ab::mbar_wait(&dkv_free, (kt_it & 1) ^ 1u, 86);   // previous key tile's dK / dV were read
                        ab::tc_fence_after();
                        const uint64_t domn = umma_desc_mn_sw128(adO + c * TILE_BYTES);
                        const uint64_t qmn = umma_desc_mn_sw128(aQ + c * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            ab::umma_ts<1>(tmem_base + COL_DV, tmem_base + COL_PT + k * 8,
                                           domn + static_cast<uint64_t>((k * 2048) >> 4), idesc_o, (c | k) != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            ab::umma_ts<1>(tmem_base + COL_DK, tmem_base + COL_DST + k * 8,
                                           qmn + static_cast<uint64_t>((k * 2048) >> 4), idesc_o, (c | k) != 0 ? 1u : 0u);
                        ab::umma_commit(&pd_free);
                        if (c == n_qc - 1) ab::umma_commit(&dkv_full);
                        if (c + 1 < n_qc) issue_sp(c + 1);
                    }
                    ab::umma_commit(&kv_empty[kb]);     // every MMA reading this K / V tile has been issued
                    ++kt_it;
                }
                ab::umma_commit(&qdo_empty);            // ... and this item's Q / dO
            }
        }
    } else if (warp == 2) {
        // ===================== per-item tables (one item ahead) =====================
        const int tid = static_cast<int>(lane);
        const bool has_bias = (p.bias_delta != nullptr) || (p.causal_value != 0.f);
        int item_it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
            const int b = item / p.H, h = item % p.H;
            const int buf = item_it & 1;
            ab::mbar_wait(&aux_empty[buf], ((item_it >> 1) & 1) ^ 1u, 87);
#pragma unroll 4
            for (int j = tid; j < lk_pad; j += AUX_THREADS)
                s_mask[buf][j] = (j < p.Lk) ? (p.add_mask ? p.add_mask[static_cast<size_t>(b) * p.Lk + j] * LOG2E : 0.f)
                                            : -INFINITY;
            const size_t si = (static_cast<size_t>(b) * p.H + h) * p.Lq;
#pragma unroll 4
            for (int i = tid; i < lq_pad; i += AUX_THREADS) {
                s_lse[buf][i] = i < p.Lq ? p.lse[si + i] * LOG2E : INFINITY;
                s_d[buf][i] = i < p.Lq ? p.dsum[si + i] : 0.f;
            }
            if (has_bias)
#pragma unroll 4
                for (int d = tid; d < 2 * MAX_L; d += AUX_THREADS) {
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
        // ===================== P^T / dS^T + output: two threads per key row =====================
        const uint32_t lg = warp & 3u;
        const uint32_t part = (warp - 3u) >> 2;              // which 64 of the chunk's 128 query columns
        const int row = static_cast<int>(lg * 32 + lane);
        const uint32_t lane_addr = tmem_base + ((lg * 32u) << 16);
        const float scale2 = p.scale * LOG2E;
        const bool has_bias = (p.bias_delta != nullptr) || (p.causal_value != 0.f);
        int item_it = 0, kt_it = 0, ch = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
            const int b = item / p.H, h = item % p.H;
            const int buf = item_it & 1;
            ab::mbar_wait(&aux_full[buf], (item_it >> 1) & 1, 88);
            const float* bias2 = s_bias[buf];
            for (int kt = 0; kt < n_kt; ++kt) {
                const int j = kt * BLOCK + row;                 // this thread's key
                const bool live = j < p.Lk;
                if (tile_dead(b, kt)) {                         // all-masked keys: dK = dV = 0 (what the full computation yields)
                    if (live) {
                        const int64_t grow = static_cast<int64_t>(b) * p.Lk + j;
                        uint4* dv_dst = reinterpret_cast<uint4*>(p.dv + grow * p.lddv + p.dv_col0 + h * D + part * 32);
                        uint4* dk_dst = reinterpret_cast<uint4*>(p.dk + grow * p.lddk + p.dk_col0 + h * D + part * 32);
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) {
                            dv_dst[v4] = make_uint4(0u, 0u, 0u, 0u);
                            dk_dst[v4] = make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                    continue;
                }
                const float mk = s_mask[buf][min(j, lk_pad - 1)];
                const int jb = j + p.Lq - 1;                    // bias index = jb - i
                for (int c = 0; c < n_qc; ++c, ++ch) {
                    ab::mbar_wait(&sp_full, ch & 1, 89);
                    ab::tc_fence_after();
                    uint32_t pkp[2][16], pkd[2][16];
#pragma unroll
                    for (int piece = 0; piece < 2; ++piece) {
                        const int col0 = static_cast<int>(part) * 64 + piece * 32;
                        const int i0 = c * BLOCK + col0;                                // query position of column 0
                        uint32_t rs[32], rd[32];
                        ab::tmem_ld32(lane_addr + col0, rs);
                        ab::tmem_ld32(lane_addr + COL_DPT + col0, rd);
                        ab::tmem_ld_wait();
                        const float4* l4 = reinterpret_cast<const float4*>(s_lse[buf] + i0);
                        const float4* d4 = reinterpret_cast<const float4*>(s_d[buf] + i0);
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {
                            const float4 lv = l4[q4], dv = d4[q4];
                            const float ls[4] = {lv.x, lv.y, lv.z, lv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
                            float pr[4], ds[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int jj = 4 * q4 + e;
                                float a = mk;
                                if (has_bias) a += bias2[min(max(jb - (i0 + jj), 0), 2 * MAX_L - 1)];
                                const float t = fmaf(__uint_as_float(rs[jj]), scale2, a);
                                pr[e] = ex2_approx(t - ls[e]);                          // 0 for padding keys / queries
                                ds[e] = pr[e] * (__uint_as_float(rd[jj]) - dd[e]);
                            }
                            pkp[piece][2 * q4] = ab::pack2_rn<kBF16>(pr[0], pr[1]);
                            pkp[piece][2 * q4 + 1] = ab::pack2_rn<kBF16>(pr[2], pr[3]);
                            pkd[piece][2 * q4] = ab::pack2_rn<kBF16>(ds[0], ds[1]);
                            pkd[piece][2 * q4 + 1] = ab::pack2_rn<kBF16>(ds[2], ds[3]);
                        }
                    }
                    ab::mbar_wait(&pd_free, (ch & 1) ^ 1u, 90);      // the TS MMAs of the previous chunk have read P^T / dS^T
                    ab::tc_fence_after();
                    tmem_st16(lane_addr + COL_PT + part * 32, pkp[0]);
                    tmem_st16(lane_addr + COL_PT + part * 32 + 16, pkp[1]);
                    tmem_st16(lane_addr + COL_DST + part * 32, pkd[0]);
                    tmem_st16(lane_addr + COL_DST + part * 32 + 16, pkd[1]);
                    ab::tmem_st_wait();
                    ab::tc_fence_before();
                    ab::mbar_arrive(&pd_ready);
                }
                // ---- this key tile's dV and dK: 32 of the 64 columns of each per thread ----
                ab::mbar_wait(&dkv_full, kt_it & 1, 91);
                ab::tc_fence_after();
                {
                    uint32_t rv[32], rk[32];
                    ab::tmem_ld32(lane_addr + COL_DV + part * 32, rv);
                    ab::tmem_ld32(lane_addr + COL_DK + part * 32, rk);
                    ab::tmem_ld_wait();
                    if (live) {
                        const int64_t grow = static_cast<int64_t>(b) * p.Lk + j;
                        uint4* dv_dst = reinterpret_cast<uint4*>(p.dv + grow * p.lddv + p.dv_col0 + h * D + part * 32);
                        uint4* dk_dst = reinterpret_cast<uint4*>(p.dk + grow * p.lddk + p.dk_col0 + h * D + part * 32);
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) {
                            dv_dst[v4] = make_uint4(
                                ab::pack2_rn<kBF16>(__uint_as_float(rv[8 * v4]), __uint_as_float(rv[8 * v4 + 1])),
                                ab::pack2_rn<kBF16>(__uint_as_float(rv[8 * v4 + 2]), __uint_as_float(rv[8 * v4 + 3])),
                                ab::pack2_rn<kBF16>(__uint_as_float(rv[8 * v4 + 4]), __uint_as_float(rv[8 * v4 + 5])),
                                ab::pack2_rn<kBF16>(__uint_as_float(rv[8 * v4 + 6]), __uint_as_float(rv[8 * v4 + 7])));
                            dk_dst[v4] = make_uint4(
                                ab::pack2_rn<kBF16>(__uint_as_float(rk[8 * v4]) * p.scale, __uint_as_float(rk[8 * v4 + 1]) * p.scale),
                                ab::pack2_rn<kBF16>(__uint_as_float(rk[8 * v4 + 2]) * p.scale, __uint_as_float(rk[8 * v4 + 3]) * p.scale),
                                ab::pack2_rn<kBF16>(__uint_as_float(rk[8 * v4 + 4]) * p.scale, __uint_as_float(rk[8 * v4 + 5]) * p.scale),
                                ab::pack2_rn<kBF16>(__uint_as_float(rk[8 * v4 + 6]) * p.scale, __uint_as_float(rk[8 * v4 + 7]) * p.scale));
                        }
                    }
                }
                ab::tc_fence_before();
                ab::mbar_arrive(&dkv_free);
                ++kt_it;
            }
            ab::mbar_arrive(&aux_empty[buf]);       // last read of this item's tables
        }
    }

    ab::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ab::tc_fence_after();
        ab::tmem_dealloc<1>(tmem_base, TMEM_COLS);
    }
}

}  // namespace attnb_tc_dkv

// Launched by atlas_b200_attention_bwd when ATLAS_B200_ATTN_BWD_TC=2 and the shape qualifies (Lq, Lk <= 512,
// Lq + Lk - 1 <= 1024).  `lse` / `dsum` must already hold the forward's log-sum-exp and the dQ kernel's D.
int atlas_b200_attn_bwd_dkv_tc(const void* q, int64_t ldq, int32_t q_col0, const void* k, int64_t ldk, int32_t k_col0,
                               const void* v, int64_t ldv, int32_t v_col0, const void* dout, int64_t lddo, void* dk,
                               int64_t lddk, int32_t dk_col0, void* dv, int64_t lddv, int32_t dv_col0, const float* add_mask,
                               const float* bias_delta, const float* lse, const float* dsum, const uint8_t* blk_live,
                               int32_t B, int32_t H, int32_t Lq, int32_t Lk, float scale, float causal_value, int32_t is_bf16,
                               cudaStream_t s) {
    using namespace attnb_tc_dkv;
    if (Lk > MAX_L || Lq > MAX_L || Lq + Lk - 1 > 2 * MAX_L) return ATLAS_B200_EUNSUPPORTED;
    CUtensorMap tq, tk, tv, tdo;
    int rc = abh::make_tmap_2d_16bit(&tq, q, static_cast<uint64_t>(B) * Lq, static_cast<uint64_t>(q_col0 + H * D),
                                     static_cast<uint64_t>(ldq), BLOCK, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tk, k, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(k_col0 + H * D),
                                 static_cast<uint64_t>(ldk), BLOCK, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tv, v, static_cast<uint64_t>(B) * Lk, static_cast<uint64_t>(v_col0 + H * D),
                                 static_cast<uint64_t>(ldv), BLOCK, D, is_bf16 != 0);
    if (rc) return rc;
    rc = abh::make_tmap_2d_16bit(&tdo, dout, static_cast<uint64_t>(B) * Lq, static_cast<uint64_t>(H * D),
                                 static_cast<uint64_t>(lddo), BLOCK, D, is_bf16 != 0);
    if (rc) return rc;
    Params p;
    p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.dk = static_cast<uint16_t*>(dk);
    p.dv = static_cast<uint16_t*>(dv);
    p.lddk = lddk; p.lddv = lddv;
    p.dk_col0 = dk_col0; p.dv_col0 = dv_col0;
    p.add_mask = add_mask;
    p.bias_delta = bias_delta;
    p.lse = lse;
    p.dsum = dsum;
    p.blk_live = blk_live;
    p.scale = scale;
    p.causal_value = causal_value;
    const int items = B * H;
    const int grid = items < abh::num_sms() ? items : abh::num_sms();
    static bool attr_set[2] = {false, false};
    if (is_bf16) {
        if (!attr_set[0]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set[0] = true;
        }
        attn_bwd_dkv_tc_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    } else {
        if (!attr_set[1]) {
            AB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            attr_set[1] = true;
        }
        attn_bwd_dkv_tc_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    }
    AB_CUDA_CHECK(cudaGetLastError());
    return ATLAS_B200_OK;
}
