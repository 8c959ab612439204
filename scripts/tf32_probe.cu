// Probe of tcgen05.mma kind::tf32 before building the tf32 / tf32x3 kernels on it:
//   (1) numerics of one 128 x N x 32 product with raw fp32 operands: does the tensor core TRUNCATE or ROUND the low 13 mantissa
//       bits?  error of the 1-pass product and of the 3-pass split  a_hi.b_hi + a_hi.b_lo + a_lo.b_hi  against fp64;
//   (2) the operand layouts the three kernels need: A K-major x B K-major (forward), A K-major x B MN-major (dgrad),
//       A MN-major x B MN-major (wgrad), all SWIZZLE_128B;
//   (3) cycles per M=128, N=256, K=8 instruction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../deep_successor_features_for_transfer_b200/csrc tf32_probe.cu -o tf32_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "tc_common.cuh"
using namespace sfgpi::tc;

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::tf32 instruction descriptor: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, majors @15/@16, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

constexpr int KT = 32;           // k extent of the test (one 128-byte swizzle span of fp32 = 4 MMAs of K = 8)

// byte offset of element (r, k) of a K-major [rows][32] fp32 tile, SWIZZLE_128B
__host__ __device__ inline uint32_t off_kmajor(int r, int k) { return r * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4; }
// byte offset of element (k, n) of an MN-major fp32 tile.  32-bit MN-major operands have ONE legal swizzled layout (CUTLASS
// sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem layout"): LayoutType 1 =
// SWIZZLE_128B_BASE32B = Swizzle<2,5,2> o ((32 elements, n),(4, k)) : ((1, LBO),(128 B, SBO)) -- 32-byte chunks of a 128-byte row
// XORed with (row mod 4); k atoms of 4 rows SBO = 512 B apart; 32-wide n blocks LBO apart.  TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__host__ __device__ inline uint32_t off_mnmajor(int k, int n, uint32_t lbo) {
    return (n >> 5) * lbo + (k >> 2) * 512 + (k & 3) * 128 + ((((n & 31) >> 3) ^ (k & 3)) << 5) + (n & 7) * 4;
}
__device__ __forceinline__ uint64_t desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)1 << 61);
}

// mode: 0 = A K-major, B K-major; 1 = A K-major, B MN-major; 2 = both MN-major.  passes: 1 (raw) or 3 (split).
// A_hi/A_lo: [128][32], B_hi/B_lo: [N][32] row-major fp32 in global; D out [128][N].
__global__ void __launch_bounds__(128, 1) probe(int N, int mode, int passes, const float *Ah, const float *Al, const float *Bh,
                                                const float *Bl, float *D, int reps, long long *cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x;
    // smem: A_hi 16K | A_lo 16K | B_hi 32K | B_lo 32K
    const uint32_t oAh = 0, oAl = 16384, oBh = 32768, oBl = 65536;
    const uint32_t lboA = 8 * 128 * (KT / 8), lboB = 8 * 128 * (KT / 8);      // MN-major: one 32-wide block = (KT/8) k-groups of 1024 B
    for (int e = tid; e < 128 * KT; e += 128) {
        const int r = e / KT, k = e % KT;
        const uint32_t o = (mode == 2) ? off_mnmajor(k, r, lboA) : off_kmajor(r, k);
        *reinterpret_cast<float *>(smem + oAh + o) = Ah[e];
        *reinterpret_cast<float *>(smem + oAl + o) = Al[e];
    }
    for (int e = tid; e < N * KT; e += 128) {
        const int n = e / KT, k = e % KT;
        const uint32_t o = (mode >= 1) ? off_mnmajor(k, n, lboB) : off_kmajor(n, k);
        *reinterpret_cast<float *>(smem + oBh + o) = Bh[e];
        *reinterpret_cast<float *>(smem + oBl + o) = Bl[e];
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (tid < 32) tmem_alloc(smem_u32(&holder), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    if (tid == 0) {
        const uint32_t idesc = idesc_tf32(128, N, mode == 2 ? 1u : 0u, mode >= 1 ? 1u : 0u);
        auto adesc = [&](uint32_t base, int k8) {
            return mode == 2 ? desc_mn_sw128_32b(sbase + base + k8 * 1024, lboA) : umma_desc_k_sw128(sbase + base + k8 * 32);
        };
        auto bdesc = [&](uint32_t base, int k8) {
            return mode >= 1 ? desc_mn_sw128_32b(sbase + base + k8 * 1024, lboB) : umma_desc_k_sw128(sbase + base + k8 * 32);
        };
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            uint32_t acc = r ? 1u : 0u;
            for (int k8 = 0; k8 < KT / 8; ++k8) {
                umma_tf32(tmem, adesc(oAh, k8), bdesc(oBh, k8), idesc, acc);
                acc = 1u;
                if (passes == 3) {
                    umma_tf32(tmem, adesc(oAh, k8), bdesc(oBl, k8), idesc, 1u);
                    umma_tf32(tmem, adesc(oAl, k8), bdesc(oBh, k8), idesc, 1u);
                }
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        cyc[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // drain: thread = lane = row
    const uint32_t t_lane = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(t_lane + c0, v);
        tmem_wait_ld();
        for (int i = 0; i < 8; ++i) D[(size_t)tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

static float tf32_rna(float x) { uint32_t b; memcpy(&b, &x, 4); b = (b + 0x1000u) & ~0x1FFFu; float y; memcpy(&y, &b, 4); return y; }
static float tf32_trunc(float x) { uint32_t b; memcpy(&b, &x, 4); b &= ~0x1FFFu; float y; memcpy(&y, &b, 4); return y; }

int main() {
    const int N = 256;
    std::vector<float> A(128 * KT), B(N * KT), Ahi(A.size()), Alo(A.size()), Bhi(B.size()), Blo(B.size()), Z(A.size() > B.size() ? A.size() : B.size(), 0.f);
    srand(7);
    for (auto &v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto &v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (size_t i = 0; i < A.size(); ++i) { Ahi[i] = tf32_rna(A[i]); Alo[i] = A[i] - Ahi[i]; }
    for (size_t i = 0; i < B.size(); ++i) { Bhi[i] = tf32_rna(B[i]); Blo[i] = B[i] - Bhi[i]; }
    float *dA, *dAl, *dB, *dBl, *dD;
    long long *dc, hc;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dAl, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dBl, B.size() * 4);
    cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dc, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    std::vector<float> D(128 * N);
    auto run = [&](int mode, int passes, const std::vector<float> &ah, const std::vector<float> &al, const std::vector<float> &bh,
                   const std::vector<float> &bl, int reps) {
        cudaMemcpy(dA, ah.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dAl, al.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, bh.data(), B.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBl, bl.data(), B.size() * 4, cudaMemcpyHostToDevice);
        probe<<<1, 128, 100 * 1024>>>(N, mode, passes, dA, dAl, dB, dBl, dD, reps, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    };
    auto err = [&](int kind) {      // kind 0: exact fp64 product of A, B; 1: tf32-rna inputs; 2: tf32-truncated inputs
        double worst = 0, scale = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int k = 0; k < KT; ++k) {
                    float a = A[r * KT + k], b = B[n * KT + k];
                    if (kind == 1) { a = tf32_rna(a); b = tf32_rna(b); }
                    if (kind == 2) { a = tf32_trunc(a); b = tf32_trunc(b); }
                    acc += (double)a * (double)b;
                }
                worst = fmax(worst, fabs(acc - (double)D[(size_t)r * N + n]));
                scale = fmax(scale, fabs(acc));
            }
        return worst / scale;
    };
    const char *names[3] = {"A K-major, B K-major (forward)", "A K-major, B MN-major (dgrad)", "A MN-major, B MN-major (wgrad)"};
    for (int mode = 0; mode < 3; ++mode) {
        run(mode, 1, A, Z, B, Z, 1);
        printf("%s\n  1 pass, raw fp32 operands : rel err vs exact %.3e | vs rna-rounded inputs %.3e | vs truncated inputs %.3e\n", names[mode],
               err(0), err(1), err(2));
        run(mode, 3, Ahi, Alo, Bhi, Blo, 1);
        printf("  3-pass split (hi = rna)   : rel err vs exact %.3e\n", err(0));
    }
    for (int passes : {1, 3}) {
        run(0, passes, Ahi, Alo, Bhi, Blo, 256);
        const double n_mma = 256.0 * (KT / 8) * passes;
        printf("timing: %d pass(es), N=256: %.1f cycles per M128 N256 K8 tf32 MMA (%lld cycles / %.0f MMAs)\n", passes, (double)hc / n_mma, hc, n_mma);
    }
    return 0;
}
