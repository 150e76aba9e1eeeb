// Device-to-device copy with TMA bulk copies only (global -> shared -> global, one elected thread per
// block, S stages of T bytes): what does the async proxy reach against SM-issued loads/stores
// (scripts/stream_copy.cu) and cudaMemcpy D2D?  1 GiB -> 1 GiB.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int T, int S>
__global__ void __launch_bounds__(32) tma_copy(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, int64_t n_tiles)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bars[S];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int64_t mine = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;      // tiles blockIdx.x, +grid, ...
    for (int64_t i = 0; i < mine + S - 1; ++i) {
        if (i < mine) {
            const int s = static_cast<int>(i % S);
            if (i >= S) bulk_wait_read0();                       // the store that read this stage has drained it
            const uint32_t bar = smem_u32(&bars[s]);
            mbar_expect_tx(bar, T);
            bulk_load(smem_u32(smem + s * T), src + (blockIdx.x + i * gridDim.x) * (int64_t)T, T, bar);
        }
        const int64_t j = i - (S - 1);
        if (j >= 0) {
            const int s = static_cast<int>(j % S);
            mbar_wait(smem_u32(&bars[s]), static_cast<uint32_t>((j / S) & 1));
            bulk_store(dst + (blockIdx.x + j * gridDim.x) * (int64_t)T, smem_u32(smem + s * T), T);
            bulk_commit();
        }
    }
    bulk_wait_all();
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

template <int T, int S> void run(int sm, int bps, const unsigned char *s, unsigned char *d, int64_t bytes)
{
    cudaFuncSetAttribute(tma_copy<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, T * S);
    float ms = time_ms([&] { tma_copy<T, S><<<sm * bps, 32, T * S>>>(s, d, bytes / T); }, 20);
    printf("tile %3d KB x %d stages, %d blocks/SM: %7.1f us  %5.0f GB/s  (%s)\n", T / 1024, S, bps, ms * 1e3, 2.0 * bytes / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int64_t bytes = 1ll << 30;
    unsigned char *s, *d; cudaMalloc(&s, bytes); cudaMalloc(&d, bytes); cudaMemset(s, 7, bytes); cudaMemset(d, 0, bytes);
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    run<16384, 4>(sm, 2, s, d, bytes);
    run<16384, 4>(sm, 3, s, d, bytes);
    run<32768, 3>(sm, 2, s, d, bytes);
    run<32768, 2>(sm, 3, s, d, bytes);
    run<8192, 4>(sm, 6, s, d, bytes);
    run<4096, 8>(sm, 6, s, d, bytes);
    run<65536, 3>(sm, 1, s, d, bytes);
    unsigned char h[4]; cudaMemcpy(h, d + bytes - 4, 4, cudaMemcpyDeviceToHost);
    printf("last bytes %d %d %d %d (expect 7)  %s\n", h[0], h[1], h[2], h[3], cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
