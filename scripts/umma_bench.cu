// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128) as a function of N, operands in smem (SS).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../deep_successor_features_for_transfer_b200/csrc umma_bench.cu -o umma_bench
#include <cstdio>
#include "tc_common.cuh"
using namespace sfgpi::tc;

__global__ void __launch_bounds__(128, 1) bench(int N, int reps, int distinct_k, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const uint32_t sbase = smem_u32(smem);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&holder), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint32_t a_base = sbase, b_base = sbase + 64 * 1024;
        long long t0 = clock64();
        if (distinct_k > 0) {
            for (int r = 0; r < reps; ++r) {
                const int kk = r % distinct_k;
                const uint64_t ad = umma_desc_k_sw128(a_base + (kk >> 2) * 16384 + (kk & 3) * 32);
                const uint64_t bd = umma_desc_k_sw128(b_base + (kk >> 2) * 32768 + (kk & 3) * 32);
                umma_bf16(tmem, ad, bd, idesc, r ? 1u : 0u);
            }
        } else {
            // optimised issue: base descriptors hoisted, k16 steps unrolled, only 64-bit adds of immediates in the loop
            const uint64_t ad0 = umma_desc_k_sw128(a_base), bd0 = umma_desc_k_sw128(b_base);
            for (int r = 0; r < reps; r += 4) {
                const uint64_t ad = ad0 + (uint64_t)(((r >> 2) & 3) * (16384 >> 4));
                const uint64_t bd = bd0 + (uint64_t)(((r >> 2) & 1) * (32768 >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, (r | k) ? 1u : 0u);
            }
        }
        umma_commit(smem_u32(&bar));
        long long t1 = clock64();
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int Ns[] = {16, 32, 64, 128, 256};
    for (int N : Ns)
        for (int reps : {64, 512}) {
            bench<<<1, 128, 200 * 1024>>>(N, reps, 16, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("N=%3d reps=%4d: issue %lld cyc, complete %lld cyc -> %.1f cyc/MMA (%s)\n", N, reps, h[0], h[1], (double)h[1] / reps,
                   cudaGetErrorString(e));
        }
    for (int N : Ns) {
        bench<<<1, 128, 200 * 1024>>>(N, 512, 0, d);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("optimised issue N=%3d: issue %lld complete %lld -> %.1f cyc/MMA\n", N, h[0], h[1], (double)h[1] / 512);
    }
    for (int N : {128, 256}) {
        bench<<<148, 128, 200 * 1024>>>(N, 512, 0, d);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("optimised issue, 148 CTAs N=%3d: %.1f cyc/MMA\n", N, (double)h[1] / 512);
    }
    // all SMs busy at once (power / shared resources)
    for (int N : {128, 256}) {
        bench<<<148, 128, 200 * 1024>>>(N, 512, 16, d);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("148 CTAs N=%3d: %.1f cyc/MMA\n", N, (double)h[1] / 512);
    }
    return 0;
}
