// Device helpers of the peer-memory exchange (csrc/peer.cu explains the protocol); shared with the TD kernel, which pulls the
// GPI keys of its rows from the peers itself.
#pragma once
#include "common.cuh"

namespace sfgpi {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// peer loads must not be served from this SM's L1 (the line may hold the previous epoch): relaxed.sys goes to the owner's L2
__device__ __forceinline__ longlong2 ld_peer_i64x2(const long long *p) {
    longlong2 v;
    asm volatile("ld.relaxed.sys.global.v2.s64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long ld_peer_i64(const long long *p) {
    long long v;
    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer_f32(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// The signalling CTA (`signaller`) stores `epoch` into every peer's flag block -- and into its own slot of the LOCAL block, which
// orders this kernel's other CTAs behind whatever the signaller did first (e.g. packing x_local); every CTA then waits until
// all ranks, this one included, have signalled `epoch` on `channel`.
__device__ __forceinline__ void peer_signal_and_wait(const sfgpi_peer_ctx &c, int channel, unsigned long long epoch, bool signaller) {
    const int t = threadIdx.x;
    if (t < c.world) {
        if (signaller) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<unsigned long long *>(c.flags[t]) + channel * SFGPI_MAX_PEERS + c.rank, epoch);
        }
        const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(c.flags[c.rank]) + channel * SFGPI_MAX_PEERS + t;
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys(mine) < epoch) {
            if (globaltimer_ns() - t0 > 20ull * 1000000000ull) {
                printf("sfgpi peer exchange: rank %d waited 20 s for rank %d (channel %d, epoch %llu) -- aborting\n", c.rank, t, channel,
                       epoch);
                __trap();
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

}  // namespace sfgpi
