// TD target, the two MSE losses and every gradient that does not flow through the psi MLP, fused in one kernel
// (sfdqn.py:330-345; tsfdqn.py:621-645; features/deep.py:112-121).  Elementwise / tiny-GEMM work: HBM-bound on the
// [B][D] operands, deterministic two-stage reductions (per-CTA partials, summed by the Adam kernel).
//
// Thread-block clusters: 8 CTAs (8 x 32 transitions) form a cluster and sum their gradient / loss partials through
// distributed shared memory in a fixed rank order before anything is written, so the Adam kernel reads ceil(B/256) partials
// per parameter instead of ceil(B/32) (that read loop was the longest latency chain of the whole train step).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace sfgpi {

constexpr int kTdRows = 32;       // transitions per CTA
constexpr int kTdThreads = 128;
constexpr int kTdCluster = 8;     // CTAs per cluster (portable maximum)

// smem layout (floats): see carve-up below
__global__ void __cluster_dims__(kTdCluster, 1, 1) __launch_bounds__(kTdThreads)
td_kernel(const __grid_constant__ sfgpi_td_args a) {
    extern __shared__ __align__(16) float sm[];
    pdl_launch_dependents();
    pdl_wait();
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int B = a.B, S = a.S, D = a.D, G = a.G;
    const int pl = blockIdx.y, blk = blockIdx.x;
    const int clu = blk / kTdCluster, nclu = gridDim.x / kTdCluster;
    const int row0 = blk * kTdRows;
    const int rows = max(0, min(kTdRows, B - row0));         // the grid is padded to whole clusters: trailing CTAs add zeros
    const bool tsf = a.variant == 2, reward = a.variant >= 1;
    const int n_acc = a.aux_len + 2;                         // [aux gradient partials | sum diff^2 | sum e^2]

    // carve-up
    float *w_s = sm;                              // [D]
    float *phi_s = w_s + D;                       // [32][D]  phi (raw)
    float *tphi_s = phi_s + kTdRows * D;          // [32][D]  phi~ (transformed)
    float *diff_s = tphi_s + kTdRows * D;         // [32][D]  then reused as daff
    float *e_s = diff_s + kTdRows * D;            // [32]
    float *red_s = e_s + kTdRows;                 // [8] block-reduction scratch
    float *Wg_s = red_s + 8;                      // [G][S]   (tsf only from here)
    float *bg_s = Wg_s + G * S;                   // [G]
    float *Wh_s = bg_s + G;                       // [D][G]
    float *bh_s = Wh_s + D * G;                   // [D]
    float *ss_s = bh_s + D;                       // [32][S]  s + s'
    float *u_s = ss_s + kTdRows * S;              // [32][G]  g(s) + g(s')
    float *du_s = u_s + kTdRows * G;              // [32][G]
    float *gacc_s = tsf ? du_s + kTdRows * G : Wg_s;      // [aux_len + 2] this CTA's partials, summed across the cluster
    for (int e = tid; e < n_acc; e += kTdThreads) gacc_s[e] = 0.0f;

    const float *cur = a.cur_sel + (size_t)pl * B * D;
    const float *nxt = a.next_sel + (size_t)pl * B * D;
    float *dout = a.d_out + (size_t)pl * B * D;

    if (reward) for (int d = tid; d < D; d += kTdThreads) w_s[d] = a.w[(size_t)pl * a.w_stride + d];
    for (int e = tid; e < kTdRows * D; e += kTdThreads) {
        int r = e / D;
        phi_s[e] = r < rows ? a.phis[(size_t)row0 * D + e] : 0.0f;
    }
    if (tsf) {
        const float *gp = a.g + (size_t)pl * a.g_stride;
        for (int e = tid; e < G * S; e += kTdThreads) Wg_s[e] = gp[e];
        for (int e = tid; e < G; e += kTdThreads) bg_s[e] = gp[G * S + e];
        for (int e = tid; e < D * G; e += kTdThreads) Wh_s[e] = a.h[e];
        for (int e = tid; e < D; e += kTdThreads) bh_s[e] = a.h[D * G + e];
        for (int e = tid; e < kTdRows * S; e += kTdThreads) {
            int r = e / S;
            ss_s[e] = r < rows ? a.states[(size_t)row0 * S + e] + a.next_states[(size_t)row0 * S + e] : 0.0f;
        }
    }
    __syncthreads();

    // u = g(s) + g(s') = Wg (s + s') + 2 bg                                   (tsfdqn.py:621-622)
    if (tsf) {
        // thread = one g column, 4 rows per trip: each weight is read once per 4 FMAs and there is no per-element division
        for (int g = tid; g < G; g += kTdThreads) {
            const float b2 = 2.0f * bg_s[g];
            const float *wg = Wg_s + g * S;
            for (int r = 0; r < kTdRows; r += 4) {
                float a0 = b2, a1 = b2, a2 = b2, a3 = b2;
                for (int s = 0; s < S; ++s) {
                    const float w = wg[s];
                    a0 = fmaf(w, ss_s[(r + 0) * S + s], a0);
                    a1 = fmaf(w, ss_s[(r + 1) * S + s], a1);
                    a2 = fmaf(w, ss_s[(r + 2) * S + s], a2);
                    a3 = fmaf(w, ss_s[(r + 3) * S + s], a3);
                }
                u_s[(r + 0) * G + g] = a0; u_s[(r + 1) * G + g] = a1; u_s[(r + 2) * G + g] = a2; u_s[(r + 3) * G + g] = a3;
            }
        }
        __syncthreads();
    }

    // phi~, target, diff, d_out, l1 partial                                    (tsfdqn.py:623-633 / sfdqn.py:330-335)
    const float c1 = 2.0f / ((float)B * (float)a.A * (float)D);
    float l1_acc = 0.0f;
    for (int e = tid; e < kTdRows * D; e += kTdThreads) {
        int r = e / D, d = e - r * D;
        float tphi = phi_s[e], df = 0.0f;
        if (tsf) {
            float aff = 2.0f * bh_s[d];
            for (int g = 0; g < G; ++g) aff = fmaf(Wh_s[d * G + g], u_s[r * G + g], aff);
            tphi *= aff;
        }
        if (r < rows) {
            const size_t gi = (size_t)row0 * D + e;
            float nv;
            if (a.next_psi != nullptr) {                 // gather psi^-(s')[a*] from the full target output by the GPI key
                const int astar = (int)key_index(a.next_keys[(size_t)pl * a.next_key_stride + row0 + r]);
                nv = a.next_psi[((size_t)(row0 + r) * a.n_pol + pl) * ((size_t)a.A * D) + (size_t)astar * D + d];
            } else {
                nv = nxt[gi];
            }
            const float target = fmaf(a.gammas[row0 + r], nv, tphi);
            df = cur[gi] - target;
            dout[gi] = c1 * df;
            l1_acc = fmaf(df, df, l1_acc);
        }
        tphi_s[e] = tphi;
        diff_s[e] = df;
    }
    __syncthreads();

    // reward head: e = w . phi~ - r, l2 partial                                (sfdqn.py:339-341 / tsfdqn.py:638-642)
    float l2_acc = 0.0f;
    if (reward) {
        if (tid < kTdRows) {
            float ev = 0.0f;
            if (tid < rows) {
                for (int d = 0; d < D; ++d) ev = fmaf(w_s[d], tphi_s[tid * D + d], ev);
                ev -= a.rs[row0 + tid];
                l2_acc = ev * ev;
            }
            e_s[tid] = ev;
        }
        __syncthreads();
    }

    // loss partials (deterministic tree inside the CTA, one slot per CTA)
    {
        float s1 = warp_sum(l1_acc), s2 = warp_sum(l2_acc);
        if ((tid & 31) == 0) { red_s[tid >> 5] = s1; red_s[4 + (tid >> 5)] = s2; }
        __syncthreads();
        if (tid == 0) {
            gacc_s[a.aux_len] = (red_s[0] + red_s[1]) + (red_s[2] + red_s[3]);
            gacc_s[a.aux_len + 1] = (red_s[4] + red_s[5]) + (red_s[6] + red_s[7]);
        }
    }
    float *gpart = gacc_s;
    const float c2 = 2.0f * a.beta / (float)B;           // variant 1: beta == 1

    if (reward) {
        // dL/dw[d] = c2 * sum_b e_b * phi~[b][d]
        for (int d = tid; d < D; d += kTdThreads) {
            float acc = 0.0f;
            for (int r = 0; r < kTdRows; ++r) acc = fmaf(e_s[r], tphi_s[r * D + d], acc);
            gpart[d] = c2 * acc;
        }
    }
    if (tsf) {

    // daff = (dL/dphi~) * phi,  dL/dphi~ = -c1*diff + c2*e*w                   (targets carry grad, tsfdqn.py:629)
    __syncthreads();
    for (int e = tid; e < kTdRows * D; e += kTdThreads) {
        int r = e / D, d = e - r * D;
        diff_s[e] = (c2 * e_s[r] * w_s[d] - c1 * diff_s[e]) * phi_s[e];
    }
    __syncthreads();
    float *daff_s = diff_s;
    // du = daff . Wh
    for (int g = tid; g < G; g += kTdThreads) {
        for (int r = 0; r < kTdRows; r += 4) {
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
            for (int d = 0; d < D; ++d) {
                const float w = Wh_s[d * G + g];
                a0 = fmaf(daff_s[(r + 0) * D + d], w, a0);
                a1 = fmaf(daff_s[(r + 1) * D + d], w, a1);
                a2 = fmaf(daff_s[(r + 2) * D + d], w, a2);
                a3 = fmaf(daff_s[(r + 3) * D + d], w, a3);
            }
            du_s[(r + 0) * G + g] = a0; du_s[(r + 1) * G + g] = a1; du_s[(r + 2) * G + g] = a2; du_s[(r + 3) * G + g] = a3;
        }
    }
    __syncthreads();
    // parameter gradients: g.W [G][S], g.b [G], h.W [D][G], h.b [D]
    float *gW = gpart + D, *gb = gW + G * S, *hW = gb + G, *hb = hW + D * G;
    // thread = one g column; 4 outputs (4 s / 4 d) per trip share each du / u load
    for (int g = tid; g < G; g += kTdThreads) {
        float accb = 0.0f;
        for (int r = 0; r < kTdRows; ++r) accb += du_s[r * G + g];
        gb[g] = 2.0f * accb;
        for (int s0 = 0; s0 < S; s0 += 4) {
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
            const int s1 = min(s0 + 1, S - 1), s2 = min(s0 + 2, S - 1), s3 = min(s0 + 3, S - 1);
            for (int r = 0; r < kTdRows; ++r) {
                const float v = du_s[r * G + g];
                const float *sr = ss_s + r * S;
                a0 = fmaf(v, sr[s0], a0); a1 = fmaf(v, sr[s1], a1); a2 = fmaf(v, sr[s2], a2); a3 = fmaf(v, sr[s3], a3);
            }
            gW[g * S + s0] = a0;
            if (s0 + 1 < S) gW[g * S + s0 + 1] = a1;
            if (s0 + 2 < S) gW[g * S + s0 + 2] = a2;
            if (s0 + 3 < S) gW[g * S + s0 + 3] = a3;
        }
        for (int d0 = 0; d0 < D; d0 += 4) {
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
            const int d1 = min(d0 + 1, D - 1), d2 = min(d0 + 2, D - 1), d3 = min(d0 + 3, D - 1);
            for (int r = 0; r < kTdRows; ++r) {
                const float v = u_s[r * G + g];
                const float *dr = daff_s + r * D;
                a0 = fmaf(dr[d0], v, a0); a1 = fmaf(dr[d1], v, a1); a2 = fmaf(dr[d2], v, a2); a3 = fmaf(dr[d3], v, a3);
            }
            hW[d0 * G + g] = a0;
            if (d0 + 1 < D) hW[(d0 + 1) * G + g] = a1;
            if (d0 + 2 < D) hW[(d0 + 2) * G + g] = a2;
            if (d0 + 3 < D) hW[(d0 + 3) * G + g] = a3;
        }
    }
    for (int d = tid; d < D; d += kTdThreads) {
        float acc = 0.0f;
        for (int r = 0; r < kTdRows; ++r) acc += daff_s[r * D + d];
        hb[d] = 2.0f * acc;
    }
    }   // tsf

    // ---- cluster reduction through distributed shared memory: rank r sums slice r of the 8 CTAs' partials, rank order ----
    cluster.sync();
    {
        const unsigned rk = cluster.block_rank();
        const int per = (n_acc + kTdCluster - 1) / kTdCluster;
        const int e_lo = rk * per, e_hi = min(n_acc, e_lo + per);
        const float *remote[kTdCluster];
#pragma unroll
        for (int q = 0; q < kTdCluster; ++q) remote[q] = cluster.map_shared_rank(gacc_s, q);
        for (int e = e_lo + tid; e < e_hi; e += kTdThreads) {
            float v[kTdCluster];
#pragma unroll
            for (int q = 0; q < kTdCluster; ++q) v[q] = remote[q][e];
            const float sum = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            if (e < a.aux_len) a.aux_grad_part[((size_t)pl * nclu + clu) * a.aux_len + e] = sum;
            else a.loss_part[((size_t)pl * nclu + clu) * 2 + (e - a.aux_len)] = sum;
        }
    }
    cluster.sync();                                          // nobody leaves while its shared memory is still being read
}

}  // namespace sfgpi

using namespace sfgpi;

extern "C" int sfgpi_td_partials(int32_t B) { return B <= 0 ? 1 : (B + kTdRows * kTdCluster - 1) / (kTdRows * kTdCluster); }

extern "C" int sfgpi_td_step(const sfgpi_td_args *args, void *stream) {
    const sfgpi_td_args &a = *args;
    if (a.variant < 0 || a.variant > 2 || a.B < 0 || a.n_pol < 0 || a.D < 1) { set_error("sfgpi_td_step: invalid arguments"); return SFGPI_E_INVALID; }
    if (a.B == 0 || a.n_pol == 0) return SFGPI_OK;
    const bool tsf = a.variant == 2;
    const int want_aux = a.variant == 0 ? 0 : (tsf ? a.D + a.G * a.S + a.G + a.D * a.G + a.D : a.D);
    if (a.aux_len < want_aux) { set_error("sfgpi_td_step: aux_len %d < %d", a.aux_len, want_aux); return SFGPI_E_INVALID; }
    size_t fl = a.D + 3 * (size_t)kTdRows * a.D + kTdRows + 8 + (size_t)a.aux_len + 2;
    if (tsf) fl += (size_t)a.G * a.S + a.G + (size_t)a.D * a.G + a.D + (size_t)kTdRows * a.S + 2 * (size_t)kTdRows * a.G;
    const size_t bytes = fl * sizeof(float);
    if (bytes > (size_t)kMaxSmem) { set_error("sfgpi_td_step: D/G too large for shared memory (%zu B)", bytes); return SFGPI_E_SMEM; }
    if (bytes > 48 * 1024) cudaFuncSetAttribute(td_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    dim3 grid(sfgpi_td_partials(a.B) * kTdCluster, a.n_pol);
    launch_pdl(td_kernel, grid, dim3(kTdThreads), bytes, (cudaStream_t)stream, a);
    return check_launch("sfgpi_td_step");
}
