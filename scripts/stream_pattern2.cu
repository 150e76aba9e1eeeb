// Access-pattern ceiling vs. access width: the wide step kernel's streams with 4, 8 or 16 envs per
// thread (32-, 64-, 128-bit row accesses), rows handled in groups of four cells, no arithmetic.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int C = 16;
template <typename V> struct W;
template <> struct W<unsigned> { static constexpr int E = 4; };
template <> struct W<uint2> { static constexpr int E = 8; };
template <> struct W<uint4> { static constexpr int E = 16; };
__device__ inline unsigned mix(unsigned a, unsigned b) { return a ^ b; }
__device__ inline uint2 mix(uint2 a, uint2 b) { return make_uint2(a.x ^ b.x, a.y ^ b.y); }
__device__ inline uint4 mix(uint4 a, uint4 b) { return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }

template <typename V, int BPS>
__global__ void __launch_bounds__(256, BPS)
pattern(const int8_t *__restrict__ act, int8_t *st, int32_t *t, float *rew, uint32_t *idx, uint8_t *f0, uint8_t *f1,
        uint8_t *f2, uint8_t *f3, int64_t n, int64_t ld)
{
    constexpr int E = W<V>::E;
    const int64_t stride = (int64_t)gridDim.x * 256 * E;
    for (int64_t e0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * E; e0 < n; e0 += stride) {
        int4 tt[E / 4];
#pragma unroll
        for (int k = 0; k < E / 4; ++k) tt[k] = __ldcs((const int4 *)(t + e0) + k);
#pragma unroll
        for (int g = 0; g < C; g += 4) {
            V s[4], a[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                s[c] = __ldcs((const V *)(st + (g + c) * ld + e0));
                a[c] = __ldcs((const V *)(act + (g + c) * ld + e0));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) __stcs((V *)(st + (g + c) * ld + e0), mix(s[c], a[c]));
        }
#pragma unroll
        for (int k = 0; k < E / 4; ++k) {
            tt[k].x += 1;
            __stcs((int4 *)(t + e0) + k, tt[k]);
            __stcs((int4 *)(rew + e0) + k, tt[k]);
            __stcs((int4 *)(idx + e0) + k, tt[k]);
        }
        V z = {};
        __stcs((V *)(f0 + e0), z); __stcs((V *)(f1 + e0), z); __stcs((V *)(f2 + e0), z); __stcs((V *)(f3 + e0), z);
    }
}

template <typename F>
float time_ms(F f, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

template <typename V, int BPS>
void run(const char *name, int sm, int8_t *act, int8_t *st, int32_t *t, float *rew, uint32_t *idx, uint8_t **f, int64_t n)
{
    float ms = time_ms([&] { pattern<V, BPS><<<sm * BPS, 256>>>(act, st, t, rew, idx, f[0], f[1], f[2], f[3], n, n); }, 50);
    printf("%s, %d blocks/SM: %.1f us  %.0f GB/s\n", name, BPS, ms * 1e3, 68.0 * n / ms / 1e6);
}

int main()
{
    const int64_t n = 1 << 24, ld = n;
    int8_t *act, *st; int32_t *t; float *rew; uint32_t *idx; uint8_t *f[4];
    cudaMalloc(&act, C * ld); cudaMalloc(&st, C * ld); cudaMalloc(&t, 4 * ld); cudaMalloc(&rew, 4 * ld); cudaMalloc(&idx, 4 * ld);
    for (auto &p : f) cudaMalloc(&p, ld);
    cudaMemset(act, 1, C * ld); cudaMemset(st, 0, C * ld); cudaMemset(t, 0, 4 * ld);
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    run<unsigned, 2>("32-bit rows (4 envs/thread)", sm, act, st, t, rew, idx, f, n);
    run<unsigned, 4>("32-bit rows (4 envs/thread)", sm, act, st, t, rew, idx, f, n);
    run<unsigned, 8>("32-bit rows (4 envs/thread)", sm, act, st, t, rew, idx, f, n);
    run<uint2, 2>("64-bit rows (8 envs/thread)", sm, act, st, t, rew, idx, f, n);
    run<uint2, 4>("64-bit rows (8 envs/thread)", sm, act, st, t, rew, idx, f, n);
    run<uint2, 8>("64-bit rows (8 envs/thread)", sm, act, st, t, rew, idx, f, n);
    run<uint4, 2>("128-bit rows (16 envs/thread)", sm, act, st, t, rew, idx, f, n);
    run<uint4, 4>("128-bit rows (16 envs/thread)", sm, act, st, t, rew, idx, f, n);
    run<uint4, 8>("128-bit rows (16 envs/thread)", sm, act, st, t, rew, idx, f, n);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
