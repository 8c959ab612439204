// Peer-memory exchange of a policy-sharded library (one process per GPU, one NVSwitch node).
//
// What the ranks exchange per train step is small and latency-bound: the packed GPI keys [n_total][B] (a MAX
// reduce-scatter: rank r needs only the rows of its own policies) and x_local = [w | delta h] (an all-gather).  As library
// collectives these are two NCCL calls that split the step's command list into three host calls and cost more than the
// kernels between them.  Here every rank maps its peers' exchange arenas (CUDA IPC) and the two exchanges become two small
// kernels INSIDE the step's kernel chain:
//   * signal: one st.release.sys per peer into the peer's flag block ("my data of epoch e is complete"),
//   * wait:   spin on the LOCAL flag block until every peer has signalled epoch e (bounded: traps after 20 s),
//   * pull:   128-bit loads straight from the peers' HBM over NVLink, reduced (MAX of keys / sum of h deltas) in registers.
// Arenas are double-buffered by epoch parity: a buffer of epoch e is rewritten at epoch e+2, after this rank has seen every
// peer's signal of epoch e+1, which a peer sends only after its pull of epoch e has finished (stream order).
#include "peer.cuh"

using namespace sfgpi;

namespace sfgpi {

// keys_out[row][b] = max_r keys_all_r[(row_lo + row)][b]: the MAX reduce-scatter of the packed GPI keys, pulled from the peers.
__global__ void __launch_bounds__(256) peer_reduce_keys_kernel(const __grid_constant__ sfgpi_peer_keys_args a) {
    pdl_launch_dependents();
    pdl_wait();                                               // the forward kernel's atomicMax results are complete and visible
    peer_signal_and_wait(a.ctx, SFGPI_PEER_CH_KEYS, (unsigned long long)a.epoch, blockIdx.x == 0);
    const int64_t n = (int64_t)a.n_rows * a.B, off = (int64_t)a.row_lo * a.B;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= n) return;
    // 128-bit pulls need every rank's element address 16-byte aligned: the element parity AND the arena base (an odd
    // n_total * B leaves the odd-epoch half of an unpadded arena only 8-byte aligned)
    bool vec = i + 1 < n;
#pragma unroll
    for (int r = 0; r < SFGPI_MAX_PEERS; ++r)
        if (r < a.ctx.world) vec = vec && ((reinterpret_cast<uintptr_t>(reinterpret_cast<const long long *>(a.keys_all[r]) + off + i) & 15) == 0);
    if (vec) {
        longlong2 v[SFGPI_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < SFGPI_MAX_PEERS; ++r)
            if (r < a.ctx.world) v[r] = ld_peer_i64x2(reinterpret_cast<const long long *>(a.keys_all[r]) + off + i);
        longlong2 m = v[0];
#pragma unroll
        for (int r = 1; r < SFGPI_MAX_PEERS; ++r)
            if (r < a.ctx.world) { m.x = max(m.x, v[r].x); m.y = max(m.y, v[r].y); }
        if ((i & 1) == 0 && (reinterpret_cast<uintptr_t>(a.keys_out) & 15) == 0) *reinterpret_cast<longlong2 *>(a.keys_out + i) = m;
        else { a.keys_out[i] = m.x; a.keys_out[i + 1] = m.y; }
    } else {
        for (int64_t k = i; k < min(i + 2, n); ++k) {
            long long m = ld_peer_i64(reinterpret_cast<const long long *>(a.keys_all[0]) + off + k);
            for (int r = 1; r < a.ctx.world; ++r) m = max(m, ld_peer_i64(reinterpret_cast<const long long *>(a.keys_all[r]) + off + k));
            a.keys_out[k] = m;
        }
    }
}

// the all-gather of x_local = [w (nw) | delta h (nh)] followed by shard_unpack_kernel's arithmetic, pulled from the peers:
// w_all [world * nw] (global policy order), h = h_prev + sum_r delta_r (rank order: identical on every rank), h_prev = h.
__global__ void __launch_bounds__(256) peer_unpack_kernel(const __grid_constant__ sfgpi_peer_unpack_args a) {
    pdl_launch_dependents(SFGPI_TR_PEER_X);
    pdl_wait(SFGPI_TR_PEER_X);
    if (blockIdx.x == 0 && a.pack_w != nullptr) {
        // fused sfgpi_shard_pack: this rank's x_local = [w | h - h_prev] is written by the signalling CTA before it signals (the
        // local flag slot keeps this kernel's other CTAs, which overwrite h and h_prev, behind it)
        float *xl = const_cast<float *>(a.x[a.ctx.rank]);
        for (int e = threadIdx.x; e < a.nw + a.nh; e += blockDim.x) xl[e] = e < a.nw ? a.pack_w[e] : a.h[e - a.nw] - a.h_prev[e - a.nw];
        __syncthreads();
    }
    peer_signal_and_wait(a.ctx, SFGPI_PEER_CH_X, (unsigned long long)a.epoch, blockIdx.x == 0);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int world = a.ctx.world;
    if (i < world * a.nw) {
        const int r = i / a.nw, k = i - r * a.nw;
        a.w_all[i] = ld_peer_f32(a.x[r] + k);
    } else if (i < world * a.nw + a.nh) {
        const int k = i - world * a.nw;
        float d[SFGPI_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < SFGPI_MAX_PEERS; ++r)
            if (r < world) d[r] = ld_peer_f32(a.x[r] + a.nw + k);
        float acc = a.h_prev[k];
#pragma unroll
        for (int r = 0; r < SFGPI_MAX_PEERS; ++r)
            if (r < world) acc += d[r];
        a.h[k] = acc;
        a.h_prev[k] = acc;
    }
    trace_exit(SFGPI_TR_PEER_X);
}

}  // namespace sfgpi

static int check_ctx(const sfgpi_peer_ctx &c, const char *who) {
    if (c.world < 1 || c.world > SFGPI_MAX_PEERS || c.rank < 0 || c.rank >= c.world) {
        set_error("%s: invalid world/rank %d/%d (at most %d peers)", who, c.world, c.rank, SFGPI_MAX_PEERS);
        return SFGPI_E_INVALID;
    }
    for (int r = 0; r < c.world; ++r)
        if (!c.flags[r]) { set_error("%s: flag block of rank %d not mapped", who, r); return SFGPI_E_INVALID; }
    return SFGPI_OK;
}

extern "C" int sfgpi_peer_reduce_keys(const sfgpi_peer_keys_args *a, void *stream) {
    if (!a) { set_error("sfgpi_peer_reduce_keys: null args"); return SFGPI_E_INVALID; }
    if (int rc = check_ctx(a->ctx, "sfgpi_peer_reduce_keys")) return rc;
    if (a->n_rows < 0 || a->B < 0 || a->row_lo < 0 || a->epoch <= 0 || !a->keys_out) {
        set_error("sfgpi_peer_reduce_keys: invalid arguments");
        return SFGPI_E_INVALID;
    }
    for (int r = 0; r < a->ctx.world; ++r)
        if (!a->keys_all[r]) { set_error("sfgpi_peer_reduce_keys: keys of rank %d not mapped", r); return SFGPI_E_INVALID; }
    const int64_t n2 = ((int64_t)a->n_rows * a->B + 1) / 2;
    launch_pdl(peer_reduce_keys_kernel, dim3((unsigned)((n2 + 255) / 256 > 0 ? (n2 + 255) / 256 : 1)), dim3(256), 0, (cudaStream_t)stream, *a);
    return check_launch("sfgpi_peer_reduce_keys");
}

extern "C" int sfgpi_peer_unpack(const sfgpi_peer_unpack_args *a, void *stream) {
    if (!a) { set_error("sfgpi_peer_unpack: null args"); return SFGPI_E_INVALID; }
    trace_bind();
    if (int rc = check_ctx(a->ctx, "sfgpi_peer_unpack")) return rc;
    if (a->nw < 0 || a->nh < 0 || a->epoch <= 0 || !a->w_all || (a->nh > 0 && (!a->h || !a->h_prev))) {
        set_error("sfgpi_peer_unpack: invalid arguments");
        return SFGPI_E_INVALID;
    }
    for (int r = 0; r < a->ctx.world; ++r)
        if (!a->x[r]) { set_error("sfgpi_peer_unpack: x_local of rank %d not mapped", r); return SFGPI_E_INVALID; }
    const int n = a->ctx.world * a->nw + a->nh;
    launch_pdl(peer_unpack_kernel, dim3((n + 255) / 256 > 0 ? (n + 255) / 256 : 1), dim3(256), 0, (cudaStream_t)stream, *a);
    return check_launch("sfgpi_peer_unpack");
}

// ---- arenas: plain cudaMalloc memory (zeroed) exported / mapped through CUDA IPC ------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == SFGPI_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int sfgpi_peer_alloc(int64_t bytes, void **dptr, void *ipc_handle_out) {
    if (bytes <= 0 || !dptr || !ipc_handle_out) { set_error("sfgpi_peer_alloc: invalid arguments"); return SFGPI_E_INVALID; }
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) return check_launch("sfgpi_peer_alloc(cudaMalloc)");
    if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        cudaFree(p);
        return check_launch("sfgpi_peer_alloc(cudaMemset)");
    }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        cudaFree(p);
        return check_launch("sfgpi_peer_alloc(cudaIpcGetMemHandle)");
    }
    memcpy(ipc_handle_out, &h, sizeof(h));
    *dptr = p;
    return SFGPI_OK;
}

extern "C" int sfgpi_peer_open(const void *ipc_handle, void **dptr) {
    if (!ipc_handle || !dptr) { set_error("sfgpi_peer_open: invalid arguments"); return SFGPI_E_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return check_launch("sfgpi_peer_open(cudaIpcOpenMemHandle)");
    *dptr = p;
    return SFGPI_OK;
}

extern "C" int sfgpi_peer_close(void *dptr) {
    if (dptr && cudaIpcCloseMemHandle(dptr) != cudaSuccess) return check_launch("sfgpi_peer_close");
    return SFGPI_OK;
}

extern "C" int sfgpi_peer_free(void *dptr) {
    if (dptr && cudaFree(dptr) != cudaSuccess) return check_launch("sfgpi_peer_free");
    return SFGPI_OK;
}
