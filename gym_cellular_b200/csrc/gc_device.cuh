// Device-side helpers shared by the kernel translation units.
#pragma once
#include "gc_internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kEPT = 4;

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, counter based: word w of env e at step t = philox(key=seed, ctr=(e_lo,e_hi,t,w/4))[w%4]
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int e) { return (w >> (8 * e)) & 0xFFu; }

// streaming accesses: every byte is used once, keep it out of L1
__device__ __forceinline__ uint32_t ld_stream_u32(const void *p)
{
    return __ldcs(reinterpret_cast<const unsigned int *>(p));
}
__device__ __forceinline__ void st_stream_u32(void *p, uint32_t v)
{
    __stcs(reinterpret_cast<unsigned int *>(p), v);
}
__device__ __forceinline__ int4 ld_stream_v4(const void *p) { return __ldcs(reinterpret_cast<const int4 *>(p)); }
__device__ __forceinline__ void st_stream_v4(void *p, int4 v) { __stcs(reinterpret_cast<int4 *>(p), v); }

// Per-thread statistics, reduced once per block at kernel exit: warp shuffles, one shared-memory
// atomic per warp, one global atomic per block and statistic.
struct ThreadStats {
    unsigned long long steps, unsafe, count, truncated;
    long long reward_q24;
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void block_flush_stats(const ThreadStats &ts, unsigned long long *s_stats,
                                                  unsigned long long *g_stats)
{
    // s_stats zeroed before the main loop (with a __syncthreads in between)
    unsigned long long v[5] = {ts.steps, ts.unsafe, ts.count, ts.truncated,
                               static_cast<unsigned long long>(ts.reward_q24)};
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const unsigned long long w = warp_sum(v[i]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_stats[i], w);
    }
    __syncthreads();
    if (threadIdx.x < 5 && s_stats[threadIdx.x]) atomicAdd(&g_stats[threadIdx.x], s_stats[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// Launch geometry: one wave of resident blocks (SM count x occupancy), grid-stride inside.
template <typename K>
int grid_for(K kernel, int64_t n_envs, int n_sm)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    const int64_t need = (n_envs + kThreads * kEPT - 1) / (kThreads * kEPT);
    const int64_t cap = static_cast<int64_t>(n_sm) * per_sm;
    return static_cast<int>(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace
