// Fused ensemble psi-MLP forward on the 5th-gen tensor cores (mode 1: bf16 operands, fp32 accumulation in TMEM).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer: streams the policy's bf16 weights (K-major, 128B-swizzled boxes of 128 rows x 64 k = 16 KB)
//               through a 4-stage mbarrier ring;
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N<=128, K=16) into TMEM, commits free the ring stages;
//   warps 2-5   epilogue group of tile slot X, warps 6-9 of tile slot Y: thread = TMEM lane = one state (row).  They
//               compute the tiny K=S input layer on CUDA cores, and after every MMA layer read the accumulator with
//               tcgen05.ld, apply bias + activation, round to bf16 and write the next layer's A operand straight into
//               shared memory in the UMMA K-major SWIZZLE_128B layout (activations never touch HBM).
// Two 128-row tiles (X, Y) of the same policy are in flight per CTA and ping-pong: while the epilogue group of X drains
// its accumulator, the tensor core runs Y's layer, so the pipe stays busy.  TMEM: 2 x 256 fp32 columns.
// The last layer is produced in chunks of <= 256 columns whose epilogue is psi store / gather / fused GPI (q = psi.w,
// running max/argmax per state in registers -- a thread owns a whole row, so no cross-thread reduction at all).
#include "tc_common.cuh"

namespace sfgpi {
namespace tc {

constexpr int kTM = 128;                 // rows per tile (UMMA M)
constexpr int kH = 256;                  // hidden width == K of every MMA layer (the shipped configs' MLP width)
constexpr int kKB = 64;                  // k elements per stage (one 128B swizzle span of bf16)
constexpr int kNKB = kH / kKB;           // 4
constexpr int kNB = 128;                 // weight rows (output columns) per stage
constexpr int kStageBytes = kNB * kKB * 2;          // 16 KB
constexpr int kNStage = 4;
constexpr int kABytes = kTM * kH * 2;               // 64 KB per tile slot
constexpr int kThreadsTc = 320;
constexpr int kMiscFloatsMax = 8192;                // W0 | b0 | hidden biases | out bias | reward vectors  (32 KB)

struct TcParams {
    sfgpi_forward_args a;
    int rows_per_policy;     // rows of the bf16 shadow per policy = Lh*256 + n3pad
    int n3pad;               // output width padded to a multiple of 16
    int Lh;                  // number of 256x256 MMA layers (n_layers - 2)
    int n_items;             // Lh + ceil(n3pad / 256)
    int tiles_per_policy, pairs_per_policy, total_pairs, paired;
    int misc_floats;
};

struct ItemInfo { int row_base, n_cols, col0, is_final; };

__device__ __forceinline__ ItemInfo item_info(const TcParams &p, int it) {
    ItemInfo r;
    if (it < p.Lh) { r.row_base = it * kH; r.n_cols = kH; r.col0 = 0; r.is_final = 0; }
    else {
        const int c = it - p.Lh;
        r.col0 = c * 256;
        r.n_cols = min(256, p.n3pad - r.col0);
        r.row_base = p.Lh * kH + r.col0;
        r.is_final = 1;
    }
    return r;
}

// byte offset of the 16-byte chunk holding columns [c, c+8) of row r inside a [128][256] bf16 K-major SW128 operand
__device__ __forceinline__ uint32_t a_chunk_off(int r, int c) {
    const int kb = c >> 6, j = (c & 63) >> 3;
    return (uint32_t)(kb * (kTM * 128) + r * 128 + ((j ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(kThreadsTc, 1)
mlp_forward_tc_kernel(const __grid_constant__ TcParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ uint8_t smem_raw[];
    const sfgpi_forward_args &a = p.a;
    const sfgpi_net_desc &net = a.net;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- carve-up (1024-byte aligned operand buffers) ----
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *A_s[2] = {base, base + kABytes};
    uint8_t *W_s = base + 2 * kABytes;
    float *misc = reinterpret_cast<float *>(W_s + kNStage * kStageBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(misc + kMiscFloatsMax);
    // barriers: [0,4) w_full, [4,8) w_empty, [8,10) slot_ready, [10,12) acc_full ; then the TMEM base holder
    const uint32_t bar0 = smem_u32(bars);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (kNStage + s); };
    auto SLOT_READY = [&](int s) { return bar0 + 8u * (2 * kNStage + s); };
    auto ACC_FULL = [&](int s) { return bar0 + 8u * (2 * kNStage + 2 + s); };
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 2 * kNStage + 4);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kNStage; ++s) { mbar_init(W_FULL(s), 1); mbar_init(W_EMPTY(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(SLOT_READY(s), 128); mbar_init(ACC_FULL(s), 1); }
        fence_mbar_init();
        tma_prefetch_desc(&tmap);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_holder), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    const int B = a.B, L = net.n_layers, A_ = net.n_actions, D = net.n_features, AD = A_ * D;
    const int S = net.dims[0];

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            uint32_t n = 0;
            for (int pair = blockIdx.x; pair < p.total_pairs; pair += gridDim.x) {
                const int pl = pair / p.pairs_per_policy, pip = pair - pl * p.pairs_per_policy;
                const int row0 = (a.policy_lo + pl) * p.rows_per_policy;
                const bool has_y = p.paired && (2 * pip + 1 < p.tiles_per_policy);
                for (int it = 0; it < p.n_items; ++it) {
                    const ItemInfo ii = item_info(p, it);
                    const int nblocks = (ii.n_cols + kNB - 1) / kNB;
                    for (int slot = 0; slot < (has_y ? 2 : 1); ++slot)
                        for (int kb = 0; kb < kNKB; ++kb)
                            for (int nb = 0; nb < nblocks; ++nb, ++n) {
                                const int s = n % kNStage;
                                mbar_wait(W_EMPTY(s), ((n / kNStage) & 1) ^ 1);
                                mbar_arrive_expect_tx(W_FULL(s), kStageBytes);
                                tma_load_2d(smem_u32(W_s + s * kStageBytes), &tmap, W_FULL(s), kb * kKB,
                                            row0 + ii.row_base + nb * kNB);
                            }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            uint32_t n = 0, ready_cnt[2] = {0, 0};
            for (int pair = blockIdx.x; pair < p.total_pairs; pair += gridDim.x) {
                const int pl = pair / p.pairs_per_policy, pip = pair - pl * p.pairs_per_policy;
                const bool has_y = p.paired && (2 * pip + 1 < p.tiles_per_policy);
                for (int it = 0; it < p.n_items; ++it) {
                    const ItemInfo ii = item_info(p, it);
                    const int nblocks = (ii.n_cols + kNB - 1) / kNB;
                    for (int slot = 0; slot < (has_y ? 2 : 1); ++slot) {
                        mbar_wait(SLOT_READY(slot), ready_cnt[slot] & 1);
                        ++ready_cnt[slot];
                        tc_fence_after();
                        const uint32_t a_base = smem_u32(A_s[slot]);
                        const uint32_t d_base = tmem_base + (uint32_t)slot * 256u;
                        for (int kb = 0; kb < kNKB; ++kb)
                            for (int nb = 0; nb < nblocks; ++nb, ++n) {
                                const int s = n % kNStage;
                                mbar_wait(W_FULL(s), (n / kNStage) & 1);
                                tc_fence_after();
                                const uint32_t ncols = (uint32_t)min(kNB, ii.n_cols - nb * kNB);
                                const uint32_t idesc = umma_idesc_bf16(kTM, ncols);
                                const uint32_t b_base = smem_u32(W_s + s * kStageBytes);
#pragma unroll
                                for (int k16 = 0; k16 < kKB / 16; ++k16) {
                                    const uint64_t adesc = umma_desc_k_sw128(a_base + kb * (kTM * 128) + k16 * 32);
                                    const uint64_t bdesc = umma_desc_k_sw128(b_base + k16 * 32);
                                    umma_bf16(d_base + nb * kNB, adesc, bdesc, idesc, (kb | k16) ? 1u : 0u);
                                }
                                umma_commit(W_EMPTY(s));             // stage free once these MMAs have read it
                            }
                        umma_commit(ACC_FULL(slot));                  // accumulator of this item complete
                    }
                }
            }
        }
    } else {
        // =========================== epilogue groups ===========================
        const int slot = (warp - 2) >> 2;                   // warps 2-5 -> X, 6-9 -> Y
        const int quad = warp & 3;                          // TMEM lane quadrant this warp may access
        const int r = quad * 32 + lane;                     // row inside the tile == TMEM lane
        const int et = threadIdx.x - 64;                    // 0..255 among epilogue threads
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)slot * 256u;
        uint8_t *Arow = A_s[slot];
        // misc layout
        float *W0_s = misc, *b0_s = W0_s + kH * S, *bh_s = b0_s + kH, *bo_s = bh_s + p.Lh * kH, *wv_s = bo_s + p.n3pad;
        const bool gpi = a.w != nullptr;
        const int nw = gpi ? (a.w_diag ? 1 : a.n_w) : 0;
        uint32_t full_cnt = 0;
        int cur_policy = -1;

        for (int pair = blockIdx.x; pair < p.total_pairs; pair += gridDim.x) {
            const int pl = pair / p.pairs_per_policy, pip = pair - pl * p.pairs_per_policy;
            const bool has_y = p.paired && (2 * pip + 1 < p.tiles_per_policy);
            const int tile = p.paired ? 2 * pip + slot : pip;
            const float *P = a.params + (size_t)(a.policy_lo + pl) * net.row_stride;

            if (pl != cur_policy) {                          // per-policy constants -> smem (all 256 epilogue threads)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int e = et; e < kH * S; e += 256) W0_s[e] = P[net.w_off[0] + e];
                for (int e = et; e < kH; e += 256) b0_s[e] = P[net.b_off[0] + e];
                for (int l = 0; l < p.Lh; ++l)
                    for (int e = et; e < kH; e += 256) bh_s[l * kH + e] = P[net.b_off[1 + l] + e];
                for (int e = et; e < p.n3pad; e += 256) bo_s[e] = e < AD ? P[net.b_off[L - 1] + e] : 0.0f;
                if (gpi) {
                    const float *wsrc = a.w_diag ? a.w + (size_t)pl * D : a.w;
                    for (int e = et; e < nw * D; e += 256) wv_s[e] = wsrc[e];
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                cur_policy = pl;
            }
            if (slot == 1 && !has_y) continue;               // (uniform per group) nothing to do for Y in this pair

            const int b = tile * kTM + r;                     // global state index of this thread's row
            const bool row_ok = b < B;

            // ---------------- input layer (K = S) on CUDA cores -> A operand ----------------
            {
                float xv[16];
#pragma unroll
                for (int s = 0; s < 16; ++s) xv[s] = (s < S && row_ok) ? a.x[(size_t)b * S + s] : 0.0f;
                float *save = (a.acts_out[0] && row_ok) ? a.acts_out[0] + ((size_t)pl * B + b) * kH : nullptr;
                const int act0 = net.acts[0];
                for (int c = 0; c < kH; c += 8) {
                    float h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float acc = b0_s[c + i];
                        const float *wr = W0_s + (c + i) * S;
#pragma unroll
                        for (int s = 0; s < 16; ++s)
                            if (s < S) acc = fmaf(wr[s], xv[s], acc);
                        h[i] = bf16_round(apply_act(acc, act0));
                    }
                    uint4 pk = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]),
                                          pack_bf16x2(h[6], h[7]));
                    *reinterpret_cast<uint4 *>(Arow + a_chunk_off(r, c)) = pk;
                    if (save) {
                        *reinterpret_cast<float4 *>(save + c) = make_float4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<float4 *>(save + c + 4) = make_float4(h[4], h[5], h[6], h[7]);
                    }
                }
                fence_proxy_async();                          // generic-proxy smem writes -> visible to the UMMA (async proxy)
                mbar_arrive(SLOT_READY(slot));
            }

            // gather selector / GPI running state for this row
            int sel_base = -1;
            if (a.sel_out != nullptr && row_ok) {
                int sidx = a.sel_actions ? (int)a.sel_actions[b]
                                         : (int)key_index(a.sel_keys[(size_t)pl * a.sel_key_stride + b]);
                sel_base = sidx * D;
            }
            float q4[4] = {0.f, 0.f, 0.f, 0.f}, best4[4];
            int besta4[4];
            int dcnt = 0, acnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) { best4[k] = -INFINITY; besta4[k] = 0; }

            for (int it = 0; it < p.n_items; ++it) {
                const ItemInfo ii = item_info(p, it);
                mbar_wait(ACC_FULL(slot), full_cnt & 1);
                ++full_cnt;
                tc_fence_after();
                if (!ii.is_final) {
                    // -------- hidden layer: bias + act, bf16, write next A operand (in place), optional save --------
                    const float *bias = bh_s + it * kH;
                    const int act = net.acts[1 + it];
                    float *save = (a.acts_out[1 + it] && row_ok) ? a.acts_out[1 + it] + ((size_t)pl * B + b) * kH : nullptr;
                    for (int c0 = 0; c0 < kH; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(t_lane + c0, v);
                        tmem_wait_ld();
                        float h[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) h[i] = bf16_round(apply_act(__uint_as_float(v[i]) + bias[c0 + i], act));
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 pk = make_uint4(pack_bf16x2(h[8 * g], h[8 * g + 1]), pack_bf16x2(h[8 * g + 2], h[8 * g + 3]),
                                                  pack_bf16x2(h[8 * g + 4], h[8 * g + 5]), pack_bf16x2(h[8 * g + 6], h[8 * g + 7]));
                            *reinterpret_cast<uint4 *>(Arow + a_chunk_off(r, c0 + 8 * g)) = pk;
                        }
                        if (save) {
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                *reinterpret_cast<float4 *>(save + c0 + 4 * g) = make_float4(h[4 * g], h[4 * g + 1], h[4 * g + 2], h[4 * g + 3]);
                        }
                    }
                    tc_fence_before();
                    fence_proxy_async();
                    mbar_arrive(SLOT_READY(slot));
                } else {
                    // -------- output layer chunk: psi store / gather / fused GPI --------
                    const int ngroups = gpi ? (nw + 3) >> 2 : 1;
                    for (int g = 0; g < ngroups; ++g) {
                        if (g > 0 || ii.col0 == 0) {
                            dcnt = 0; acnt = ii.col0 / D;     // (multi-group is only used when the whole layer is one chunk)
#pragma unroll
                            for (int k = 0; k < 4; ++k) { q4[k] = 0.f; best4[k] = -INFINITY; besta4[k] = 0; }
                        }
                        for (int c0 = 0; c0 < ii.n_cols; c0 += 32) {
                            uint32_t v[32];
                            tmem_ld32(t_lane + c0, v);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int col = ii.col0 + c0 + i;
                                if (col < AD && c0 + i < ii.n_cols) {
                                    const float val = __uint_as_float(v[i]) + bo_s[col];
                                    if (g == 0) {
                                        if (a.psi_out != nullptr && row_ok)
                                            a.psi_out[((size_t)b * a.n_pol + pl) * AD + col] = val;
                                        const unsigned off = (unsigned)(col - sel_base);
                                        if (sel_base >= 0 && off < (unsigned)D) a.sel_out[((size_t)pl * B + b) * D + off] = val;
                                    }
                                    if (gpi) {
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            if (g * 4 + k < nw) q4[k] = fmaf(val, wv_s[(g * 4 + k) * D + dcnt], q4[k]);
                                        if (++dcnt == D) {
#pragma unroll
                                            for (int k = 0; k < 4; ++k) {
                                                if (g * 4 + k < nw) {
                                                    if (k == 0 && g == 0 && a.q_out != nullptr && row_ok)
                                                        a.q_out[((size_t)b * a.n_pol + pl) * A_ + acnt] = q4[k];
                                                    if (q4[k] > best4[k]) { best4[k] = q4[k]; besta4[k] = acnt; }
                                                    q4[k] = 0.f;
                                                }
                                            }
                                            dcnt = 0; ++acnt;
                                        }
                                    }
                                }
                            }
                        }
                        if (gpi && it == p.n_items - 1 && row_ok) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (g * 4 + k < nw) {
                                    const int wi = a.w_diag ? pl : g * 4 + k;
                                    if (a.key_action)
                                        atomicMax(reinterpret_cast<long long *>(a.key_action) + (size_t)wi * B + b,
                                                  pack_key(best4[k], (uint32_t)besta4[k]));
                                    if (a.key_task)
                                        atomicMax(reinterpret_cast<long long *>(a.key_task) + (size_t)wi * B + b,
                                                  pack_key(best4[k], (uint32_t)(a.task_base + pl)));
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    if (it + 1 < p.n_items) mbar_arrive(SLOT_READY(slot));     // next chunk may overwrite the accumulator
                }
            }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// fp32 library rows -> bf16 shadow [n_pol][rows_per_policy][256]: rows = hidden W_1..W_Lh then W_out padded to n3pad rows
__global__ void pack_bf16_kernel(sfgpi_net_desc net, const float *__restrict__ params, int policy_lo, int n_pol,
                                 __nv_bfloat16 *__restrict__ out, int rows_per_policy, int Lh, int n3pad) {
    const int AD = net.n_actions * net.n_features, L = net.n_layers;
    const size_t total = (size_t)n_pol * rows_per_policy * kH;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % kH);
        const size_t rr = i / kH;
        const int row = (int)(rr % rows_per_policy), pl = (int)(rr / rows_per_policy);
        const float *P = params + (size_t)(policy_lo + pl) * net.row_stride;
        float v = 0.0f;
        if (row < Lh * kH) {
            const int l = 1 + row / kH, n = row % kH;
            v = P[net.w_off[l] + n * kH + k];
        } else {
            const int n = row - Lh * kH;
            if (n < AD) v = P[net.w_off[L - 1] + n * kH + k];
        }
        out[((size_t)(policy_lo + pl) * rows_per_policy + row) * kH + k] = __float2bfloat16_rn(v);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

static bool tc_shape_ok(const sfgpi_net_desc &net, const char **why) {
    if (net.n_layers < 3) { *why = "needs >= 3 Linear layers"; return false; }
    for (int l = 1; l < net.n_layers; ++l)
        if (net.dims[l] != kH) { *why = "every hidden width must be 256"; return false; }
    if (net.dims[0] > 16) { *why = "state dimension must be <= 16"; return false; }
    if (net.acts[net.n_layers - 1] != SFGPI_ACT_NONE) { *why = "output layer must be linear"; return false; }
    return true;
}

}  // namespace tc
}  // namespace sfgpi

using namespace sfgpi;
using namespace sfgpi::tc;

extern "C" int sfgpi_bf16_rows_per_policy(const sfgpi_net_desc *net) {
    const int AD = net->n_actions * net->n_features;
    return (net->n_layers - 2) * kH + ((AD + 15) & ~15);
}

extern "C" int sfgpi_pack_bf16(const sfgpi_net_desc *net, const float *params, int32_t policy_lo, int32_t n_pol, void *out_bf16,
                               void *stream) {
    const char *why = "";
    if (!tc_shape_ok(*net, &why)) { set_error("sfgpi_pack_bf16: tensor-core path %s", why); return SFGPI_E_INVALID; }
    if (n_pol <= 0) return SFGPI_OK;
    const int AD = net->n_actions * net->n_features, Lh = net->n_layers - 2, n3pad = (AD + 15) & ~15;
    const int rpp = Lh * kH + n3pad;
    const size_t total = (size_t)n_pol * rpp * kH;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*net, params, policy_lo, n_pol,
                                                               reinterpret_cast<__nv_bfloat16 *>(out_bf16), rpp, Lh, n3pad);
    return check_launch("sfgpi_pack_bf16");
}

// mode-1 forward.  `params_bf16` = shadow produced by sfgpi_pack_bf16 for the same `params`; n_rows_total = number of
// policies in the shadow (for the tensor map extent).
extern "C" int sfgpi_mlp_forward_tc(const sfgpi_forward_args *args, const void *params_bf16, int32_t n_policies_total,
                                    void *stream) {
    const sfgpi_forward_args &a = *args;
    const sfgpi_net_desc &net = a.net;
    const char *why = "";
    if (!tc_shape_ok(net, &why)) { set_error("sfgpi_mlp_forward_tc: tensor-core path %s", why); return SFGPI_E_INVALID; }
    if (a.B < 0 || a.n_pol < 0 || net.dims[net.n_layers] != net.n_actions * net.n_features) {
        set_error("sfgpi_mlp_forward_tc: invalid sizes");
        return SFGPI_E_INVALID;
    }
    if (a.B == 0 || a.n_pol == 0) return SFGPI_OK;
    TcParams p;
    p.a = a;
    const int AD = net.n_actions * net.n_features;
    p.Lh = net.n_layers - 2;
    p.n3pad = (AD + 15) & ~15;
    p.rows_per_policy = p.Lh * kH + p.n3pad;
    const int n_chunks = (p.n3pad + 255) / 256;
    p.n_items = p.Lh + n_chunks;
    const bool gpi = a.w != nullptr;
    const int nw = gpi ? (a.w_diag ? 1 : a.n_w) : 0;
    if (gpi && nw > 4 && n_chunks > 1) {
        set_error("sfgpi_mlp_forward_tc: more than 4 reward vectors need A*D <= 256 in the fused GPI epilogue");
        return SFGPI_E_INVALID;
    }
    p.misc_floats = kH * net.dims[0] + kH + p.Lh * kH + p.n3pad + nw * net.n_features;
    if (p.misc_floats > kMiscFloatsMax) {
        set_error("sfgpi_mlp_forward_tc: per-policy constants (%d floats) exceed the shared-memory budget", p.misc_floats);
        return SFGPI_E_SMEM;
    }
    p.tiles_per_policy = (a.B + kTM - 1) / kTM;
    const int total_tiles = p.tiles_per_policy * a.n_pol;
    p.paired = total_tiles > 148 ? 1 : 0;                    // small problems: one tile per CTA, no ping-pong partner
    p.pairs_per_policy = p.paired ? (p.tiles_per_policy + 1) / 2 : p.tiles_per_policy;
    p.total_pairs = p.pairs_per_policy * a.n_pol;

    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { set_error("sfgpi_mlp_forward_tc: cuTensorMapEncodeTiled entry point not found"); return SFGPI_E_CUDA; }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)kH, (cuuint64_t)n_policies_total * p.rows_per_policy};
    const cuuint64_t gstride[1] = {(cuuint64_t)kH * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kKB, (cuuint32_t)kNB};
    const cuuint32_t estride[2] = {1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(params_bf16), gdim, gstride, box, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("sfgpi_mlp_forward_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr); return SFGPI_E_CUDA; }

    const int smem_bytes = 1024 + 2 * kABytes + kNStage * kStageBytes + kMiscFloatsMax * 4 + 256;
    static bool cfg = false;
    if (!cfg) { cudaFuncSetAttribute(mlp_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); cfg = true; }
    const int grid = p.total_pairs < 148 ? p.total_pairs : 148;
    mlp_forward_tc_kernel<<<grid, kThreadsTc, smem_bytes, (cudaStream_t)stream>>>(p, tmap);
    return check_launch("sfgpi_mlp_forward_tc");
}
