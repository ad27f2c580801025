// Microbenchmark: cycles per tcgen05.mma (kind::f16, K=16) as a function of N, A source (smem / TMEM)
// and cta_group.  Operands are whatever is in smem/TMEM (values irrelevant).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I atlas_b200/csrc -o gpurun_out/umma_bench tools/umma_bench.cu
#include "common.cuh"
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t rnd(uint32_t& st) { st = st * 1664525u + 1013904223u; return st; }

template <int CG, bool TS>
__global__ void __launch_bounds__(128, 1) bench(int n_dim, int iters, long long* out, int randomize) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_base_smem;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ab::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t rank = CG == 2 ? ab::cluster_ctarank() : 0;
    uint32_t seed = threadIdx.x * 7919u + blockIdx.x * 104729u + 1u;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) {
        // random fp16 pairs in [-2, 2): sign/exponent/mantissa all toggling (worst case for datapath power)
        uint32_t a = rnd(seed), b = rnd(seed);
        uint32_t h0 = (a & 0x8000u) | (0x3800u + (a & 0x07FFu)), h1 = (b & 0x8000u) | (0x3800u + (b & 0x07FFu));
        reinterpret_cast<uint32_t*>(smem_raw)[i] = randomize ? (h0 | (h1 << 16)) : 0u;
    }
    if (warp == 0 && lane == 0) { ab::mbar_init(&done_bar, 1); ab::fence_barrier_init(); }
    if (warp == 1) ab::tmem_alloc<CG>(&tmem_base_smem, 512);
    ab::fence_proxy_async_smem();
    ab::tc_fence_before();
    if (CG == 2) ab::cluster_sync_all(); else __syncthreads();
    ab::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;
    if (randomize) {  // fill the A region of TMEM (columns [0,256)) with random fp16 pairs
        uint32_t r[32];
        for (int c = 0; c < 8; ++c) {
            for (int j = 0; j < 32; ++j) { uint32_t a = rnd(seed), b = rnd(seed);
                r[j] = ((a & 0x8000u) | (0x3800u + (a & 0x07FFu))) | (((b & 0x8000u) | (0x3800u + (b & 0x07FFu))) << 16); }
            ab::tmem_st32(tmem + ((warp * 32u) << 16) + c * 32, r);
        }
        ab::tmem_st_wait();
        ab::tc_fence_before();
    }
    __syncthreads();
    ab::tc_fence_after();
    if (warp == 0 && lane == 0 && rank == 0) {
        const uint32_t idesc = ab::umma_idesc_f16(128 * CG, n_dim, false);
        const uint64_t adesc = ab::umma_desc_k_sw128(smem_base);
        const uint64_t bdesc = ab::umma_desc_k_sw128(smem_base + 16384);
        const uint32_t d = tmem + 256;   // accumulators in columns [256, 512)
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t off = ((i & 7) * 4096) >> 4;  // walk over 8 different 4 KB slabs like the real kernels
                if (TS) ab::umma_ts<CG>(d, tmem + ((i & 7) * 32) + (k * 8), bdesc + off + 2 * k, idesc, (i | k) ? 1u : 0u);
                else    ab::umma_ss<CG>(d, adesc + 2 * k, bdesc + off + 2 * k, idesc, (i | k) ? 1u : 0u);
            }
        }
        long long t1 = clock64();
        if (CG == 2) ab::umma_commit_2sm(&done_bar, 0x1); else ab::umma_commit(&done_bar);
        ab::mbar_wait(&done_bar, 0, 99);
        long long t2 = clock64();
        out[2 * blockIdx.x + 0] = t1 - t0;
        out[2 * blockIdx.x + 1] = t2 - t0;
    }
    ab::tc_fence_before();
    if (CG == 2) ab::cluster_sync_all(); else __syncthreads();
    if (warp == 1) { ab::tc_fence_after(); ab::tmem_dealloc<CG>(tmem, 512); }
}

template <int CG, bool TS>
void run(int n_dim, long long* dout, int ctas, int randomize) {
    const int iters = 4096;
    cudaFuncSetAttribute(bench<CG, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 100 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaLaunchKernelEx(&cfg, bench<CG, TS>, n_dim, iters, dout, randomize);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    }
    std::vector<long long> hv(2 * ctas);
    cudaMemcpy(hv.data(), dout, sizeof(long long) * 2 * ctas, cudaMemcpyDeviceToHost);
    long long h[2] = {0, 0};
    int cnt = 0;
    for (int b = 0; b < ctas; b += CG) { h[0] += hv[2 * b]; h[1] += hv[2 * b + 1]; ++cnt; }
    h[0] /= cnt; h[1] /= cnt;
    const double n_mma = iters * 4.0;
    const double macs = 128.0 * CG * n_dim * 16;
    printf("ctas %3d %s  cta_group %d  A-from-%s  M=%3d N=%3d : issue %.1f cyc/mma, complete %.1f cyc/mma  -> %.0f MAC/cyc/SM\n", ctas, randomize ? "random" : "zeros ", CG,
           TS ? "TMEM" : "smem", 128 * CG, n_dim, h[0] / n_mma, h[1] / n_mma, macs / (h[1] / n_mma) / CG);
}

int main() {
    long long* dout;
    cudaMalloc(&dout, 16 * 148);
    cudaMemset(dout, 0, 16 * 148);
    for (int randomize : {0, 1}) {
        for (int ctas : {2, 148}) {
            for (int n : {64, 128, 256}) {
                run<1, false>(n, dout, ctas, randomize);
                run<1, true>(n, dout, ctas, randomize);
                run<2, true>(n, dout, ctas, randomize);
            }
        }
    }
    return 0;
}
