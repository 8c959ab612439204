// G4 joint psi / phi step (features/deep_phi.py:95-224): TD target with the LEARNED features phi, both losses, the learnable
// loss coefficient, and every gradient that does not pass through an MLP -- one CTA, deterministic reductions.  See
// include/sfgpi.h (sfgpi_g4_args) for the mathematics.  The phi agents run at batch 32: this step is launch-latency-bound, the
// kernel exists to keep the whole update on the device and free of host round trips (the reference issues ~100 eager ops).
#include "common.cuh"

namespace sfgpi {

constexpr int kG4Threads = 256;
constexpr int kG4MaxB = 4096;

__global__ void __launch_bounds__(kG4Threads) g4_head_kernel(const __grid_constant__ sfgpi_g4_args a) {
    __shared__ float e_s[kG4MaxB];
    __shared__ float red_s[3][kG4Threads / 32];
    __shared__ float tot_s[3];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, B = a.B, D = a.D;
    const float coef = a.coef[0], bias = a.bias[0];
    const float c1 = coef * 2.0f / ((float)B * (float)a.A * (float)D);      // d(c * psi_loss)/d cur
    const float c2 = 2.0f / (float)B;
    float l1 = 0.0f, l2 = 0.0f, se = 0.0f;
    for (int b = tid; b < B; b += kG4Threads) {
        const float *phi = a.phi + (size_t)b * D, *cur = a.cur_sel + (size_t)b * D, *nxt = a.next_sel + (size_t)b * D;
        const float gm = a.gammas[b];
        float rfit = bias;
        for (int d = 0; d < D; ++d) rfit = fmaf(a.w[d], phi[d], rfit);
        const float e = rfit - a.rs[b];
        e_s[b] = e;
        l2 = fmaf(e, e, l2);
        se += e;
        for (int d = 0; d < D; ++d) {
            const float diff = cur[d] - fmaf(gm, nxt[d], phi[d]);           // cur - targets, targets = phi + gamma * psi^-(s')[a*]
            l1 = fmaf(diff, diff, l1);
            a.d_psi[(size_t)b * D + d] = c1 * diff;
            a.d_phi[(size_t)b * D + d] = fmaf(c2 * e, a.w[d], -c1 * diff);  // phi enters the targets (sign -) and the reward fit
        }
    }
    l1 = warp_sum(l1); l2 = warp_sum(l2); se = warp_sum(se);
    if ((tid & 31) == 0) { red_s[0][tid >> 5] = l1; red_s[1][tid >> 5] = l2; red_s[2][tid >> 5] = se; }
    __syncthreads();
    if (tid < 3) {
        float t = 0.0f;
        for (int k = 0; k < kG4Threads / 32; ++k) t += red_s[tid][k];
        tot_s[tid] = t;
    }
    __syncthreads();
    const float psi_loss = tot_s[0] / ((float)B * (float)a.A * (float)D), phi_loss = tot_s[1] / (float)B;
    // dLoss/dw[d] = (2 / B) sum_b e_b phi[b][d]   (fixed order)
    for (int d = tid; d < D; d += kG4Threads) {
        float acc = 0.0f;
        for (int b = 0; b < B; ++b) acc = fmaf(e_s[b], a.phi[(size_t)b * D + d], acc);
        a.grad_small[d] = c2 * acc;
    }
    if (tid == 0) {
        a.grad_small[D] = c2 * tot_s[2];                                    // dLoss/db
        a.grad_small[D + 1] = -psi_loss;                                    // maximize=True: Adam sees -dLoss/dc
        a.losses[0] = phi_loss + coef * psi_loss;
        a.losses[1] = psi_loss;
        a.losses[2] = phi_loss;
        a.losses[3] = coef;
    }
}

}  // namespace sfgpi

using namespace sfgpi;

extern "C" int sfgpi_g4_head(const sfgpi_g4_args *args, void *stream) {
    if (!args) { set_error("sfgpi_g4_head: null args"); return SFGPI_E_INVALID; }
    const sfgpi_g4_args &a = *args;
    if (a.B < 1 || a.B > kG4MaxB || a.A < 1 || a.D < 1) { set_error("sfgpi_g4_head: needs 1 <= B <= %d, A, D >= 1", kG4MaxB); return SFGPI_E_INVALID; }
    if (!a.cur_sel || !a.next_sel || !a.phi || !a.rs || !a.gammas || !a.w || !a.bias || !a.coef || !a.d_psi || !a.d_phi || !a.grad_small || !a.losses) {
        set_error("sfgpi_g4_head: null buffer");
        return SFGPI_E_INVALID;
    }
    launch_pdl(g4_head_kernel, dim3(1), dim3(kG4Threads), 0, (cudaStream_t)stream, a);
    return check_launch("sfgpi_g4_head");
}
