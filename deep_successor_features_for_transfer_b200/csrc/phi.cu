// Reward-regression head of the learned-feature (phi) path: the loss of SFDQN.pre_train (sfdqn_phi.py:850-862)
//     phis = phi_theta(cat[s, a, s'])  [B][D]      lin_b = w_i . phis_b      loss = mean_b (r_b - lin_b)^2
// and what autograd hands back to the two optimizers:  dL/dphis[b][d] = (2/B) e_b w[d],  dL/dw[d] = (2/B) sum_b e_b phis[b][d]
// with e_b = lin_b - r_b.  The phi MLP itself runs on the ensemble kernels (forward: sfgpi_mlp_forward with A = 1; backward:
// sfgpi_mlp_backward fed with d_out from here), Adam on sfgpi_adam_step, which also reduces the loss partials written here.
// Tiny and latency-bound (B = 32 in the reference): one CTA per 256 transitions, fixed-order reductions (deterministic).
#include "common.cuh"

namespace sfgpi {

constexpr int kPhiThreads = 256;

__global__ void __launch_bounds__(kPhiThreads) phi_head_kernel(const float *__restrict__ phi, const float *__restrict__ w,
                                                               const float *__restrict__ r, int B, int D, float *__restrict__ d_out,
                                                               float *__restrict__ dw_part, float *__restrict__ loss_part) {
    __shared__ float e_s[kPhiThreads];
    __shared__ float red_s[kPhiThreads / 32];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, b = blockIdx.x * kPhiThreads + tid;
    const float c = 2.0f / (float)B;
    float e = 0.0f;
    if (b < B) {
        const float *pr = phi + (size_t)b * D;
        float lin = 0.0f;
        for (int d = 0; d < D; ++d) lin = fmaf(pr[d], w[d], lin);
        e = lin - r[b];
        float *dr = d_out + (size_t)b * D;
        for (int d = 0; d < D; ++d) dr[d] = c * e * w[d];
    }
    e_s[tid] = e;
    float sq = warp_sum(e * e);
    if ((tid & 31) == 0) red_s[tid >> 5] = sq;
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        for (int k = 0; k < kPhiThreads / 32; ++k) s += red_s[k];
        loss_part[2 * blockIdx.x] = 0.0f;                   // (l1 slot of sfgpi_adam_step's loss reduction: unused here)
        loss_part[2 * blockIdx.x + 1] = s;
    }
    const int rows = min(kPhiThreads, B - blockIdx.x * kPhiThreads);
    for (int d = tid; d < D; d += kPhiThreads) {            // dw[d] partial of this CTA, rows in ascending order
        float acc = 0.0f;
        const float *pc = phi + (size_t)blockIdx.x * kPhiThreads * D + d;
        for (int k = 0; k < rows; ++k) acc = fmaf(e_s[k], pc[(size_t)k * D], acc);
        dw_part[(size_t)blockIdx.x * D + d] = c * acc;
    }
}

}  // namespace sfgpi

using namespace sfgpi;

extern "C" int sfgpi_phi_head_partials(int32_t B) { return B <= 0 ? 1 : (B + kPhiThreads - 1) / kPhiThreads; }

extern "C" int sfgpi_phi_head(const float *phi, const float *w, const float *r, int32_t B, int32_t D, float *d_out, float *dw_part,
                              float *loss_part, void *stream) {
    if (B < 0 || D < 1 || !phi || !w || !r || !d_out || !dw_part || !loss_part) { set_error("sfgpi_phi_head: invalid arguments"); return SFGPI_E_INVALID; }
    if (B == 0) return SFGPI_OK;
    launch_pdl(phi_head_kernel, dim3(sfgpi_phi_head_partials(B)), dim3(kPhiThreads), 0, (cudaStream_t)stream, phi, w, r, (int)B, (int)D, d_out,
               dw_part, loss_part);
    return check_launch("sfgpi_phi_head");
}
