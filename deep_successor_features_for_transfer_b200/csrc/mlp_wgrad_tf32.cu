// wgrad of the psi MLP on tcgen05 kind::tf32 (one pass or the 3-pass hi / lo split), the fp32-precision twin of
// mlp_wgrad_tc_kernel (mlp_backward_tc.cu):  dW_l[n][k] = sum_b dZ_l[b][n] in_l[b][k],  db_l[n] = sum_b dZ_l[b][n], split-K over
// the batch.  Both operands are MN-major (the reduction index b is the ROW of the row-major [B][width] tensors the streaming
// forward / dgrad kernels stored): fp32 MN-major has one legal swizzled layout, SWIZZLE_128B_BASE32B, which TMA produces with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B from {32 columns x 32 rows} boxes (scripts/tf32_probe.cu verified all three layouts).
// One CTA = (policy, layer, 128-row tile of dW_l, batch split): acc[128][256] in TMEM plus 32 auxiliary columns against
// xo = [x | 1 | 0..] that yield the bias gradient (the ones column) and, for layer 0, dW_0 itself.  A ring stage holds 32 batch
// rows: dZ tile 4 boxes + input tile 8 boxes + xo 1 box per part (52 KB; hi + lo: 104 KB, 2 stages).
#include "stream_tc.cuh"
#include <stdlib.h>

namespace sfgpi {
namespace tc {

constexpr int kWfThreads = 288;                     // 4 producer warps + 1 MMA warp + 4 epilogue warps
constexpr int kWfBox = 32 * 128;                    // 4 KB: 32 batch rows x 32 fp32
constexpr int kWfA = 4 * kWfBox, kWfB = 8 * kWfBox, kWfX = kWfBox;
constexpr int kWfPart = kWfA + kWfB + kWfX;         // 52 KB

struct WfParams {
    sfgpi_net_desc net;
    int n_pol, B, L, AD, S;
    int n_split, bs;
    int mt_out, items_per_policy;
    float *grad_part;                // [n_pol][n_split][row_stride]
};

template <int PARTS>
__global__ void __launch_bounds__(kWfThreads, 1)
mlp_wgrad_tf32_kernel(const __grid_constant__ WfParams p, const __grid_constant__ CUtensorMap tmap_dz, const __grid_constant__ CUtensorMap tmap_dzo,
                      const __grid_constant__ CUtensorMap tmap_acts, const __grid_constant__ CUtensorMap tmap_xo) {
    constexpr int NS = PARTS == 2 ? 2 : 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    pdl_launch_dependents();
    const sfgpi_net_desc &net = p.net;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem_raw);
    const uint32_t bar0 = sbase + NS * PARTS * kWfPart;
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (4 + s); };
    const uint32_t ACC_FULL = bar0 + 8u * 8;
    const uint32_t holder_addr = ACC_FULL + 8u;

    const int pl = blockIdx.y;
    const int split = blockIdx.x / p.items_per_policy;
    int t = blockIdx.x - split * p.items_per_policy;
    int l, mt;
    if (t < p.mt_out) { l = p.L - 1; mt = t; }
    else { t -= p.mt_out; l = p.L - 2 - (t >> 1); mt = t & 1; }
    const bool main_mma = l >= 1;
    const int b_lo = split * p.bs, b_hi = min(p.B, b_lo + p.bs);
    const int n_kb = (b_hi - b_lo + 31) / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(FULL(s), 4); mbar_init(EMPTY(s), 1); }
        mbar_init(ACC_FULL, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmap_dz); tma_prefetch_desc(&tmap_dzo); tma_prefetch_desc(&tmap_acts); tma_prefetch_desc(&tmap_xo);
    }
    if (warp == 4) tmem_alloc(holder_addr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(holder_addr));
    pdl_wait();

    if (warp < 4) {
        // producers: the 13 boxes of a part are dealt round-robin to the 4 warps; every warp arrives once per stage
        const uint32_t leader = elect_one();
        const int n_box = (main_mma ? 13 : 5);                       // per part: 4 dZ + (8 input) + 1 xo
        int mine = 0;
        for (int q = warp; q < n_box * PARTS; q += 4) ++mine;
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % NS;
            mbar_wait_warp(EMPTY(s), ((kb / NS) & 1) ^ 1);
            mbar_arrive_expect_tx_e(FULL(s), (uint32_t)mine * kWfBox, leader);
            const int b0 = b_lo + kb * 32;
            for (int q = warp; q < n_box * PARTS; q += 4) {
                const int part = q / n_box;
                int j = q - part * n_box;
                const uint32_t st = sbase + (s * PARTS + part) * kWfPart;
                if (j < 4) {                                             // dZ tile: columns mt*128 + 32 j
                    if (l == p.L - 1) tma_load_3d_e(st + j * kWfBox, &tmap_dzo, FULL(s), mt * 128 + j * 32, b0, part * p.n_pol + pl, leader);
                    else tma_load_3d_e(st + j * kWfBox, &tmap_dz, FULL(s), mt * 128 + j * 32, b0, (part * (p.L - 1) + l) * p.n_pol + pl, leader);
                } else if (main_mma && j < 12) {                         // layer input: columns 32 (j - 4)
                    j -= 4;
                    tma_load_3d_e(st + kWfA + j * kWfBox, &tmap_acts, FULL(s), j * 32, b0, (part * (p.L - 1) + (l - 1)) * p.n_pol + pl, leader);
                } else {
                    tma_load_3d_e(st + kWfA + kWfB, &tmap_xo, FULL(s), 0, b0, part, leader);
                }
            }
        }
    } else if (warp == 4) {
        const uint32_t leader = elect_one();
        const uint32_t idesc_main = idesc_tf32(kTM, 256, 1u, 1u), idesc_aux = idesc_tf32(kTM, 32, 1u, 1u);
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % NS;
            mbar_wait_warp(FULL(s), (kb / NS) & 1);
            tc_fence_after();
            const uint32_t sh = sbase + (s * PARTS) * kWfPart, sl = sh + kWfPart;
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
                const uint32_t first = (kb | k8) ? 1u : 0u;
                const uint64_t ah = umma_desc_mn_sw128_32b(sh + k8 * 1024, kWfBox);
                const uint64_t bh = umma_desc_mn_sw128_32b(sh + kWfA + k8 * 1024, kWfBox);
                const uint64_t xh = umma_desc_mn_sw128_32b(sh + kWfA + kWfB + k8 * 1024, kWfBox);
                if (PARTS == 2) {
                    const uint64_t al = umma_desc_mn_sw128_32b(sl + k8 * 1024, kWfBox);
                    const uint64_t bl = umma_desc_mn_sw128_32b(sl + kWfA + k8 * 1024, kWfBox);
                    const uint64_t xl = umma_desc_mn_sw128_32b(sl + kWfA + kWfB + k8 * 1024, kWfBox);
                    if (main_mma) {
                        umma_tf32_e(tmem_base, al, bh, idesc_main, first, leader);
                        umma_tf32_e(tmem_base, ah, bl, idesc_main, 1u, leader);
                        umma_tf32_e(tmem_base, ah, bh, idesc_main, 1u, leader);
                    }
                    umma_tf32_e(tmem_base + 256u, al, xh, idesc_aux, first, leader);
                    umma_tf32_e(tmem_base + 256u, ah, xl, idesc_aux, 1u, leader);
                    umma_tf32_e(tmem_base + 256u, ah, xh, idesc_aux, 1u, leader);
                } else {
                    if (main_mma) umma_tf32_e(tmem_base, ah, bh, idesc_main, first, leader);
                    umma_tf32_e(tmem_base + 256u, ah, xh, idesc_aux, first, leader);
                }
            }
            umma_commit_e(EMPTY(s), leader);
        }
        umma_commit_e(ACC_FULL, leader);
    } else {
        // ---- epilogue: thread = TMEM lane = one row n of dW_l ----
        const int quad = warp & 3;
        const int n_loc = quad * 32 + lane;
        const int n = mt * 128 + n_loc;
        const int N_l = net.dims[l + 1], K_l = net.dims[l];
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        float *gp = p.grad_part + ((size_t)pl * p.n_split + split) * net.row_stride;
        mbar_wait(ACC_FULL, 0);
        tc_fence_after();
        const bool ok = n < N_l;
        if (main_mma) {
            float *wrow = gp + net.w_off[l] + (size_t)n * K_l;              // K_l == 256
#pragma unroll 1
            for (int c0 = 0; c0 < kH; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(t_lane + c0, v);
                tmem_wait_ld();
                if (ok) {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        *reinterpret_cast<float4 *>(wrow + c0 + 4 * g) =
                            make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                        __uint_as_float(v[4 * g + 3]));
                }
            }
        }
        // aux columns: [0,S) = sum_b dZ[b][n] x[b][s] (dW_0 when l == 0), column S = sum_b dZ[b][n] (bias gradient)
        {
            uint32_t v[32];
            tmem_ld32(t_lane + 256u, v);
            tmem_wait_ld();
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i == p.S) gp[net.b_off[l] + n] = __uint_as_float(v[i]);
                    else if (i < p.S && l == 0) gp[net.w_off[0] + (size_t)n * p.S + i] = __uint_as_float(v[i]);
                }
            }
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc
}  // namespace sfgpi

using namespace sfgpi;
using namespace sfgpi::tc;

// launched by sfgpi_mlp_backward_stream (mlp_stream_tc.cu)
int sfgpi_wgrad_tf32_launch(const sfgpi_net_desc &net, int parts, int n_pol, int B, const float *dz, const float *dzo, const float *acts,
                            const float *xo, int ADp, float *grad_part, int n_split, cudaStream_t st) {
    const int L = net.n_layers, AD = net.n_actions * net.n_features, S = net.dims[0];
    WfParams wp;
    wp.net = net;
    wp.n_pol = n_pol; wp.B = B; wp.L = L; wp.AD = AD; wp.S = S;
    wp.n_split = n_split;
    wp.bs = (((B + n_split - 1) / n_split) + 63) & ~63;
    wp.mt_out = (AD + 127) / 128;
    wp.items_per_policy = wp.mt_out + 2 * (L - 1);
    wp.grad_part = grad_part;
    CUtensorMap tm_dz, tm_dzo, tm_acts, tm_xo;
    int rc;
    {
        const uint64_t d3[3] = {(uint64_t)kH, (uint64_t)B, (uint64_t)parts * (L - 1) * n_pol};
        const uint64_t do3[3] = {(uint64_t)ADp, (uint64_t)B, (uint64_t)parts * n_pol};
        const uint64_t dx[3] = {32u, (uint64_t)B, (uint64_t)parts};
        const uint32_t box3[3] = {32, 32, 1};
        if ((rc = make_tmap_f32(&tm_dz, dz, 3, d3, box3, true))) return rc;
        if ((rc = make_tmap_f32(&tm_dzo, dzo, 3, do3, box3, true))) return rc;
        if ((rc = make_tmap_f32(&tm_acts, acts, 3, d3, box3, true))) return rc;
        if ((rc = make_tmap_f32(&tm_xo, xo, 3, dx, box3, true))) return rc;
    }
    const int NS = parts == 2 ? 2 : 4;
    const int smem = NS * parts * kWfPart + 256;
    static bool cfg = false;
    if (!cfg) {
        cudaFuncSetAttribute(mlp_wgrad_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(mlp_wgrad_tf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cfg = true;
    }
    dim3 grid(wp.items_per_policy * n_split, n_pol);
    if (parts == 2) launch_pdl(mlp_wgrad_tf32_kernel<2>, grid, dim3(kWfThreads), smem, st, wp, tm_dz, tm_dzo, tm_acts, tm_xo);
    else launch_pdl(mlp_wgrad_tf32_kernel<1>, grid, dim3(kWfThreads), smem, st, wp, tm_dz, tm_dzo, tm_acts, tm_xo);
    return check_launch("sfgpi_mlp_backward_stream(wgrad)");
}
