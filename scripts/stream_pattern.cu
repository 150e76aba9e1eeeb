// What can the access pattern of the wide step kernel reach on its own?  Same streams as
// cell_pair_kernel<16,0,0> (16 state rows r/w, 16 action rows r, t r/w, reward/index w, 4 flag rows w;
// 68 bytes per env), same thread mapping and grid, but no arithmetic.   nvcc -arch=sm_100a -O3, run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int C = 16;
template <int BLOCKS_PER_SM>
__global__ void __launch_bounds__(256, BLOCKS_PER_SM)
pattern(const int8_t *__restrict__ act, int8_t *st, int32_t *t, float *rew, uint32_t *idx, uint8_t *f0, uint8_t *f1,
        uint8_t *f2, uint8_t *f3, int64_t n, int64_t ld)
{
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    for (int64_t e0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; e0 < n; e0 += stride) {
        uint32_t s[C], a[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            s[c] = __ldcs((const unsigned *)(st + c * ld + e0));
            a[c] = __ldcs((const unsigned *)(act + c * ld + e0));
        }
        int4 tt = __ldcs((const int4 *)(t + e0));
        uint32_t x = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) { s[c] ^= a[c]; x += s[c]; }
#pragma unroll
        for (int c = 0; c < C; ++c) __stcs((unsigned *)(st + c * ld + e0), s[c]);
        tt.x += 1; tt.y += 1; tt.z += 1; tt.w += 1;
        __stcs((int4 *)(t + e0), tt);
        __stcs((int4 *)(rew + e0), make_int4(x, x, x, x));
        __stcs((int4 *)(idx + e0), make_int4(x, x, x, x));
        __stcs((unsigned *)(f0 + e0), 0u); __stcs((unsigned *)(f1 + e0), x); __stcs((unsigned *)(f2 + e0), x); __stcs((unsigned *)(f3 + e0), x);
    }
}

__global__ void copy_kernel(const int4 *__restrict__ src, int4 *dst, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
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

int main()
{
    const int64_t n = 1 << 24, ld = n;
    int8_t *act, *st; int32_t *t; float *rew; uint32_t *idx; uint8_t *f[4];
    cudaMalloc(&act, C * ld); cudaMalloc(&st, C * ld); cudaMalloc(&t, 4 * ld); cudaMalloc(&rew, 4 * ld); cudaMalloc(&idx, 4 * ld);
    for (auto &p : f) cudaMalloc(&p, ld);
    cudaMemset(act, 1, C * ld); cudaMemset(st, 0, C * ld); cudaMemset(t, 0, 4 * ld);
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const double bytes = 68.0 * n;
    for (int bps : {2, 4, 6, 8}) {
        float ms = 0;
        auto run = [&](auto kern, int g) { ms = time_ms([&] { kern<<<g, 256>>>(act, st, t, rew, idx, f[0], f[1], f[2], f[3], n, ld); }, 50); };
        if (bps == 2) run(pattern<2>, sm * 2); else if (bps == 4) run(pattern<4>, sm * 4); else if (bps == 6) run(pattern<6>, sm * 6); else run(pattern<8>, sm * 8);
        printf("pattern, %d blocks/SM: %.1f us  %.0f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
    }
    int4 *a4, *b4; const int64_t m = (int64_t)1 << 26;            // 1 GiB each
    cudaMalloc(&a4, m * 16); cudaMalloc(&b4, m * 16);
    float ms = time_ms([&] { copy_kernel<<<sm * 8, 256>>>(a4, b4, m); }, 20);
    printf("int4 copy kernel 1 GiB -> 1 GiB: %.1f us  %.0f GB/s (read+write)\n", ms * 1e3, 2.0 * m * 16 / ms / 1e6);
    ms = time_ms([&] { cudaMemcpyAsync(b4, a4, m * 16, cudaMemcpyDeviceToDevice); }, 20);
    printf("cudaMemcpy D2D 1 GiB: %.1f us  %.0f GB/s (read+write)\n", ms * 1e3, 2.0 * m * 16 / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
