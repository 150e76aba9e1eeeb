// Plain device-to-device copy kernels with different cache hints / unrolling: what can SM-issued
// loads and stores reach against cudaMemcpy D2D on this part?  (1 GiB -> 1 GiB)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { PLAIN, CS, NC_NOALLOC, L2_256, EVICT_FIRST };

template <int MODE> __device__ inline int4 ld(const int4 *p)
{
    int4 v;
    if (MODE == CS) return __ldcs(p);
    if (MODE == NC_NOALLOC) { asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
    if (MODE == L2_256) { asm volatile("ld.global.L2::256B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
    if (MODE == EVICT_FIRST) { asm volatile("ld.global.L1::evict_first.L2::256B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
    return *p;
}
template <int MODE> __device__ inline void st(int4 *p, int4 v)
{
    if (MODE == PLAIN || MODE == L2_256) *p = v; else __stcs(p, v);
}

template <int MODE, int U, int BPS>
__global__ void __launch_bounds__(256, BPS) copyk(const int4 *__restrict__ src, int4 *__restrict__ dst, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * 256 * U;
    for (int64_t i = (int64_t)blockIdx.x * 256 * U + threadIdx.x; i < n; i += stride) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * 256 < n) v[u] = ld<MODE>(src + i + u * 256);
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * 256 < n) st<MODE>(dst + i + u * 256, v[u]);
    }
}

template <typename F> float time_ms(F f, int reps)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

template <int MODE, int U, int BPS> void run(const char *name, int sm, const int4 *s, int4 *d, int64_t n)
{
    float ms = time_ms([&] { copyk<MODE, U, BPS><<<sm * BPS, 256>>>(s, d, n); }, 20);
    printf("%-28s unroll %d, %d blocks/SM: %7.1f us  %5.0f GB/s\n", name, U, BPS, ms * 1e3, 2.0 * n * 16 / ms / 1e6);
}

int main()
{
    const int64_t bytes = 1ll << 30, n = bytes / 16;
    int4 *s, *d; cudaMalloc(&s, bytes); cudaMalloc(&d, bytes); cudaMemset(s, 1, bytes);
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    float ms = time_ms([&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice); }, 20);
    printf("cudaMemcpy D2D: %.1f us %.0f GB/s\n", ms * 1e3, 2.0 * bytes / ms / 1e6);
    run<PLAIN, 1, 8>("plain", sm, s, d, n);
    run<PLAIN, 4, 4>("plain", sm, s, d, n);
    run<PLAIN, 4, 8>("plain", sm, s, d, n);
    run<PLAIN, 8, 4>("plain", sm, s, d, n);
    run<CS, 1, 8>("ld.cs / st.cs", sm, s, d, n);
    run<CS, 4, 4>("ld.cs / st.cs", sm, s, d, n);
    run<CS, 4, 8>("ld.cs / st.cs", sm, s, d, n);
    run<CS, 8, 4>("ld.cs / st.cs", sm, s, d, n);
    run<NC_NOALLOC, 4, 4>("ld.nc.no_allocate / st.cs", sm, s, d, n);
    run<NC_NOALLOC, 4, 8>("ld.nc.no_allocate / st.cs", sm, s, d, n);
    run<L2_256, 4, 4>("ld.L2::256B / st", sm, s, d, n);
    run<L2_256, 4, 8>("ld.L2::256B / st", sm, s, d, n);
    run<EVICT_FIRST, 4, 4>("ld.evict_first.256B / st.cs", sm, s, d, n);
    run<EVICT_FIRST, 4, 8>("ld.evict_first.256B / st.cs", sm, s, d, n);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
