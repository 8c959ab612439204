// Micro-benchmark: latency of one / several in-flight TMA 2-D box loads (128 rows x 128 B, SWIZZLE_128B) from L2-resident data.
#include <cstdio>
#include <cuda.h>
#include "tc_common.cuh"
using namespace sfgpi::tc;

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(128, 1) bench1d(const uint8_t *w, int depth, int iters, int box_bytes, long long total_bytes, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    const uint32_t sbase = smem_u32(smem);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar[i]), 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        int issued = 0, done = 0;
        long long off = ((long long)blockIdx.x * 977 * box_bytes) % (total_bytes - box_bytes);
        while (done < iters) {
            while (issued < iters && issued - done < depth) {
                const int s = issued % depth;
                mbar_arrive_expect_tx(smem_u32(&bar[s]), box_bytes);
                bulk_load_1d(sbase + s * box_bytes, w + off, box_bytes, smem_u32(&bar[s]));
                off = (off + box_bytes) % (total_bytes - box_bytes);
                ++issued;
            }
            mbar_wait(smem_u32(&bar[done % depth]), (done / depth) & 1);
            ++done;
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

__global__ void __launch_bounds__(256, 1) bench1d_multi(const uint8_t *w, int nwarps, int depth, int iters, int box_bytes, long long total_bytes, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[64];
    __shared__ long long tmax[8];
    const uint32_t sbase = smem_u32(smem);
    if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(smem_u32(&bar[i]), 1); fence_mbar_init(); }
    __syncthreads();
    const int wp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && wp < nwarps) {
        long long t0 = clock64();
        int issued = 0, done = 0;
        long long off = ((long long)(blockIdx.x * 8 + wp) * 977 * box_bytes) % (total_bytes - box_bytes);
        const uint32_t mybase = sbase + wp * depth * box_bytes;
        while (done < iters) {
            while (issued < iters && issued - done < depth) {
                const int s = issued % depth;
                mbar_arrive_expect_tx(smem_u32(&bar[wp * 8 + s]), box_bytes);
                bulk_load_1d(mybase + s * box_bytes, w + off, box_bytes, smem_u32(&bar[wp * 8 + s]));
                off = (off + box_bytes) % (total_bytes - box_bytes);
                ++issued;
            }
            mbar_wait(smem_u32(&bar[wp * 8 + done % depth]), (done / depth) & 1);
            ++done;
        }
        tmax[wp] = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x == 0) { long long m = 0; for (int i = 0; i < nwarps; ++i) m = tmax[i] > m ? tmax[i] : m; out[blockIdx.x] = m; }
}

__global__ void __launch_bounds__(128, 1) bench(const __grid_constant__ CUtensorMap tmap, int depth, int iters, int rows_total, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    const uint32_t sbase = smem_u32(smem);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar[i]), 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        int issued = 0, done = 0;
        uint32_t row = (blockIdx.x * 977) % (rows_total - 128);
        while (done < iters) {
            while (issued < iters && issued - done < depth) {
                const int s = issued % depth;
                mbar_arrive_expect_tx(smem_u32(&bar[s]), 16384);
                tma_load_2d(sbase + s * 16384, &tmap, smem_u32(&bar[s]), (issued & 3) * 64, row);
                row = (row + 128) % (rows_total - 128);
                ++issued;
            }
            mbar_wait(smem_u32(&bar[done % depth]), (done / depth) & 1);
            ++done;
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int rows = 8192;       // 8192 x 256 bf16 = 4 MB: L2 resident
    void *w; cudaMalloc(&w, (size_t)rows * 512); cudaMemset(w, 0, (size_t)rows * 512);
    long long *d, h[148]; cudaMalloc(&d, 148 * 8);
    void *sym; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t gdim[2] = {256, (cuuint64_t)rows}, gstr[1] = {512};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    ((EncodeTiledFn)sym)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384);
    for (int grid : {1, 148})
        for (int depth : {1, 2, 3, 4, 6, 8}) {
            const int iters = 256;
            bench<<<grid, 128, 8 * 16384>>>(tm, depth, iters, rows, d);   // warm L2
            bench<<<grid, 128, 8 * 16384>>>(tm, depth, iters, rows, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("grid %3d depth %d: %.0f cyc per 16KB box (%.1f B/cyc/SM)  %s\n", grid, depth, (double)mx / iters, 16384.0 * iters / mx,
                   cudaGetErrorString(e));
        }
    cudaFuncSetAttribute(bench1d, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 1024);
    for (int grid : {1, 148})
        for (int box : {16384, 32768})
            for (int depth : {1, 2, 4}) {
                const int iters = 256;
                bench1d<<<grid, 128, 4 * 32768>>>((const uint8_t *)w, depth, iters, box, (long long)rows * 512, d);
                bench1d<<<grid, 128, 4 * 32768>>>((const uint8_t *)w, depth, iters, box, (long long)rows * 512, d);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("1D bulk grid %3d box %5d depth %d: %.0f cyc per box (%.1f B/cyc/SM)  %s\n", grid, box, depth, (double)mx / iters,
                       (double)box * iters / mx, cudaGetErrorString(e));
            }
    cudaFuncSetAttribute(bench1d, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid : {1, 148})
        for (int box : {65536}) {
            const int iters = 128, depth = 2;
            bench1d<<<grid, 128, 2 * 65536>>>((const uint8_t *)w, depth, iters, box, (long long)rows * 512, d);
            bench1d<<<grid, 128, 2 * 65536>>>((const uint8_t *)w, depth, iters, box, (long long)rows * 512, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("1D bulk grid %3d box %5d depth %d: %.0f cyc per box (%.1f B/cyc/SM)  %s\n", grid, box, depth, (double)mx / iters,
                   (double)box * iters / mx, cudaGetErrorString(e));
        }
    cudaFuncSetAttribute(bench1d_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid : {1, 148})
        for (int nw : {1, 2, 4, 6}) {
            const int iters = 256, depth = 2, box = 16384;
            bench1d_multi<<<grid, 256, nw * depth * box>>>((const uint8_t *)w, nw, depth, iters, box, (long long)rows * 512, d);
            bench1d_multi<<<grid, 256, nw * depth * box>>>((const uint8_t *)w, nw, depth, iters, box, (long long)rows * 512, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("1D bulk multi-issuer grid %3d warps %d box 16K depth 2: %.1f B/cyc/SM  %s\n", grid, nw,
                   (double)box * iters * nw / mx, cudaGetErrorString(e));
        }
    return 0;
}
