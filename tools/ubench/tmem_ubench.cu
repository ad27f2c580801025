// Micro-benchmark of tensor-memory reads / writes (tcgen05.ld / tcgen05.st, 32x32b.x32) as the softmax warps of the
// attention kernels issue them: bytes per clock per SM as a function of the number of resident warps, with and without a
// dependent MUFU / FMA body.  Answers: is reading S out of TMEM (128 x Lk fp32 per query tile) a bound next to the MUFU pipe?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../atlas_b200/csrc -o tmem_ubench tmem_ubench.cu -lcuda
#include "common.cuh"

#include <cstdio>
#include <cuda_runtime.h>

// MODE: 0 = ld only, 1 = ld + 32 ex2 per load, 2 = st only, 3 = ld + st of 16 packed words (the softmax round trip, no math)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters, long long* clk_out) {
    __shared__ uint32_t tmem_base_smem;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) ab::tmem_alloc<1>(&tmem_base_smem, 512);
    ab::tc_fence_before();
    __syncthreads();
    ab::tc_fence_after();
    const uint32_t base = tmem_base_smem + (((warp & 3u) * 32u) << 16);
    const uint32_t group = warp >> 2;                  // warps of one group cover the 128 TMEM lanes
    uint32_t r[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) r[e] = threadIdx.x + e;
    float acc = 0.f;
    // initialise the columns this warp will read
    for (int c = 0; c < 512; c += 32) ab::tmem_st32(base + c, r);
    ab::tmem_st_wait();
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = ((it + group * 3) * 32) & 511u;
        if (MODE == 0 || MODE == 1 || MODE == 3) {
            ab::tmem_ld32(base + col, r);
            ab::tmem_ld_wait();
        }
        if (MODE == 1) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                float y;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(r[e]) * 1e-9f));
                acc += y;
            }
        } else if (MODE == 0 || MODE == 3) {
#pragma unroll
            for (int e = 0; e < 32; e += 8) acc += __uint_as_float(r[e]);
        }
        if (MODE == 2) {
            r[it & 31] += it;
            ab::tmem_st32(base + col, r);
            ab::tmem_st_wait();
        }
        if (MODE == 3) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = r[2 * e] ^ r[2 * e + 1];
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(base + col),
                "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]), "r"(pk[8]),
                "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
                : "memory");
            ab::tmem_st_wait();
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk_out = t1 - t0;
    __syncthreads();
    if (warp == 0) ab::tmem_dealloc<1>(tmem_base_smem, 512);
}

template <int MODE>
void run(const char* name, int warps, float* out, long long* clk) {
    const int iters = 4000;
    k<MODE><<<148, warps * 32>>>(out, 10, clk);
    k<MODE><<<148, warps * 32>>>(out, iters, clk);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double bytes = double(warps) * 32 * 32 * 4 * iters;     // fp32 words moved by the loads (or stores) per SM
    printf("%-28s warps/SM %2d: %9lld clk -> %.1f B/clk/SM (%.2f fp32 elements/clk/SM)  %s\n", name, warps, c, bytes / c,
           bytes / c / 4, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    float* out;
    long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&clk, 8);
    for (int warps : {4, 8, 12, 16}) run<0>("tcgen05.ld x32 only", warps, out, clk);
    for (int warps : {4, 8, 12, 16}) run<1>("ld x32 + 32 ex2", warps, out, clk);
    for (int warps : {4, 12}) run<2>("tcgen05.st x32 only", warps, out, clk);
    for (int warps : {4, 8, 12, 16}) run<3>("ld x32 + st x16 round trip", warps, out, clk);
    return 0;
}
