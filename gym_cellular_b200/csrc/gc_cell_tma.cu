// EXPERIMENTAL, opt-in (GC_B200_TMA=1): deterministic cellular step with TMA bulk staging for wide envs
// (C >= 8).  Bit-identical to the default kernel (tests/test_gpu_parity.py::test_tma_variant) but measured
// slower on B200 -- 0.79 (two stages, 2 blocks/SM) and 0.85 (one stage, 4 blocks/SM) of the measured HBM
// peak against 0.90 for the register-staged kernel: a tile needs 56 bulk operations of only 1 KB each and
// the two resident blocks supply too few warps for the arithmetic.  Kept as the starting point for a
// 2-D tensor-map version (one operation per [C x 1 KB] box).
//
// Same arithmetic as gc_cell_fast.cu (pair table, PRMT row rebuild, base-S^4 index folding), different
// data movement.  A block owns tiles of kTile = 1024 consecutive envs.  One elected thread moves whole
// tiles with 1-D bulk copies (cp.async.bulk, the TMA engine): the 2C int8 rows (1 KB each) and the int32
// episode steps (4 KB) of tile i+1 are in flight into one shared-memory stage while the block computes
// tile i from the other stage; completion is signalled on an mbarrier (complete_tx::bytes).  Results are
// written back to shared memory -- the next state in place over the rows it came from, t / reward /
// index / flags into an output buffer -- and leave with bulk stores (cp.async.bulk.global.shared).
//
// Why: a wide env reads 2C + 4 bytes and writes C + 16 per step through ~57 narrow memory instructions
// per thread, each with its own 64-bit address arithmetic, and the bytes a thread can keep in flight are
// bounded by its registers.  With bulk copies the bytes in flight per SM are bounded by shared memory
// (2 blocks x 36 KB at C = 16), loads never stall a warp, and the per-thread memory instructions become
// conflict-free LDS/STS with immediate offsets.
#include "gc_device.cuh"

namespace {

constexpr int kTile = kThreads * kEPT;          // envs per tile (1024)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef GC_TMA_STAGES
#define GC_TMA_STAGES 1
#endif
constexpr int kStages = GC_TMA_STAGES;          // input stages per block (1: latency hidden across blocks only)

template <int C>
struct TmaLayout {
    static constexpr int kInBytes = (2 * C + 4) * 1024;          // state rows, action rows, t
    static constexpr int kOutBytes = 16 * 1024;                  // t, reward, index (4 KB each), 4 flag rows (1 KB each)
    static constexpr int kLutBytes = (256 + 16) * 8;
    static constexpr int kTotal = kStages * kInBytes + kOutBytes + kLutBytes + 64;
};

template <int C>
__global__ void __launch_bounds__(kThreads, kStages == 1 ? 4 : 2)
cell_tma_kernel(const __grid_constant__ CellTables tab, const __grid_constant__ StepIO io, const uint2 *__restrict__ lut)
{
    using L = TmaLayout<C>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *s_in = smem;                                    // [2][kInBytes]
    unsigned char *s_out = smem + kStages * L::kInBytes;           // [kOutBytes]
    uint2 *s_pair = reinterpret_cast<uint2 *>(s_out + L::kOutBytes);
    uint2 *s_single = s_pair + 256;
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_single + 16);   // full[2]
    __shared__ unsigned long long s_stats[5];

    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += kThreads) s_pair[i] = lut[i];
    if (tid < 16) s_single[tid] = lut[GC_PAIR_LUT_PAIRS + tid];
    if (tid < 5) s_stats[tid] = 0;
    const uint32_t bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t ld = io.ld;
    const int64_t n_tiles = (io.end - io.begin + kTile - 1) / kTile;

    // producer: all rows of one tile into stage s
    auto issue_loads = [&](int64_t tile, int s) {
        const int64_t e0 = io.begin + tile * kTile;
        const int64_t left = io.end - e0;                                       // envs of this tile in range
        const uint32_t nb = static_cast<uint32_t>(((left < kTile ? left : kTile) + 15) / 16 * 16);   // bytes per int8 row
        const uint32_t bar = s ? bar1 : bar0;
        const uint32_t base = smem_u32(s_in + s * L::kInBytes);
        mbar_expect_tx(bar, nb * (2 * C + 4));
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            bulk_load(base + c * 1024, io.state + c * ld + e0, nb, bar);
            bulk_load(base + (C + c) * 1024, io.actions + c * ld + e0, nb, bar);
        }
        bulk_load(base + 2 * C * 1024, io.t + e0, nb * 4, bar);
    };

    uint32_t st_steps = 0, st_unsafe = 0, st_count = 0, st_trunc = 0;
    long long st_reward = 0;
    if (kStages == 2 && tid == 0 && blockIdx.x < n_tiles) issue_loads(blockIdx.x, 0);

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = kStages == 2 ? (it & 1) : 0;
        const int64_t e_tile = io.begin + tile * kTile;
        const int64_t e0 = e_tile + tid * kEPT;
        // The other stage and the output buffer were last read by the bulk stores of the previous tile.
        if (tid == 0) {
            bulk_wait_read0();
            if (kStages == 2) { if (tile + gridDim.x < n_tiles) issue_loads(tile + gridDim.x, s ^ 1); }
            else issue_loads(tile, 0);
        }
        __syncthreads();                                            // output buffer is free again
        mbar_wait(s ? bar1 : bar0, static_cast<uint32_t>(kStages == 2 ? ((it >> 1) & 1) : (it & 1)));

        uint32_t *rows_s = reinterpret_cast<uint32_t *>(s_in + s * L::kInBytes);      // [C][256] state words
        const uint32_t *rows_a = rows_s + C * kThreads;                                // [C][256] action words
        const int4 *tin_v = reinterpret_cast<const int4 *>(s_in + s * L::kInBytes + 2 * C * 1024);
        int4 *o_t = reinterpret_cast<int4 *>(s_out);
        int4 *o_reward = o_t + kThreads;
        int4 *o_index = o_reward + kThreads;
        uint32_t *o_flags = reinterpret_cast<uint32_t *>(o_index + kThreads);          // [4][256]: term, trunc, unsafe, count

        if (e0 < io.end) {
            const int rem = static_cast<int>(io.end - e0 < kEPT ? io.end - e0 : kEPT);
            const int4 t4 = tin_v[tid];
            int tn[kEPT] = {t4.x + 1, t4.y + 1, t4.z + 1, t4.w + 1};
            uint32_t trunc_w = 0, keep = 0xFFFFFFFFu;
            if (io.max_episode_steps > 0) {
#pragma unroll
                for (int e = 0; e < kEPT; ++e)
                    if (tn[e] >= io.max_episode_steps) { tn[e] = 0; trunc_w |= 1u << (8 * e); keep &= ~(0xFFu << (8 * e)); }
            }
            float r[kEPT] = {0.f, 0.f, 0.f, 0.f};
            uint32_t add[kEPT] = {0, 0, 0, 0}, orr[kEPT] = {0, 0, 0, 0}, first[kEPT] = {0, 0, 0, 0};
            uint32_t idx[kEPT] = {0, 0, 0, 0};
#pragma unroll
            for (int g = 0; g < (C + 3) / 4; ++g) {
                uint32_t q = 0;
#pragma unroll
                for (int i = 0; i < 4; i += 2) {
                    const int c = 4 * g + i, d = c + 1;
                    if (c >= C) break;
                    const bool pair = d < C;
                    uint32_t inf[kEPT];
                    const uint32_t sc = rows_s[c * kThreads + tid] & 0x03030303u, ac = rows_a[c * kThreads + tid] & 0x03030303u;
                    if (pair) {
                        const uint32_t sd = rows_s[d * kThreads + tid] & 0x03030303u, ad = rows_a[d * kThreads + tid] & 0x03030303u;
                        const uint32_t pidx = (ad * 4u + sd) * 16u + ac * 4u + sc;
#pragma unroll
                        for (int e = 0; e < kEPT; ++e) {
                            const uint2 ent = s_pair[byte_of(pidx, e)];
                            r[e] += __uint_as_float(ent.y);
                            inf[e] = ent.x;
                        }
                    } else {
                        const uint32_t sidx = ac * 4u + sc;
#pragma unroll
                        for (int e = 0; e < kEPT; ++e) {
                            const uint2 ent = s_single[byte_of(sidx, e) & 15u];
                            r[e] += __uint_as_float(ent.y);
                            inf[e] = ent.x;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < kEPT; ++e) {
                        add[e] += inf[e];
                        if (c == 0) first[e] = inf[e]; else orr[e] |= inf[e];
                    }
                    const uint32_t u = prmt(inf[0], inf[1], 0x7362), v = prmt(inf[2], inf[3], 0x7362);
                    const uint32_t out_c = (prmt(u, v, 0x5410) & keep) | ((0x01010101u * static_cast<uint8_t>(tab.init[c])) & ~keep);
                    rows_s[c * kThreads + tid] = out_c;                     // next state, in place
                    q += out_c * tab.place4[i];
                    if (pair) {
                        const uint32_t out_d = (prmt(u, v, 0x7632) & keep) | ((0x01010101u * static_cast<uint8_t>(tab.init[d])) & ~keep);
                        rows_s[d * kThreads + tid] = out_d;
                        q += out_d * tab.place4[i + 1];
                    }
                }
                const uint32_t place = tab.place[4 * g];
#pragma unroll
                for (int e = 0; e < kEPT; ++e) idx[e] += byte_of(q, e) * place;
            }
            uint32_t unsafe_w = 0, count_w = 0;
            float rout[kEPT];
#pragma unroll
            for (int e = 0; e < kEPT; ++e) {
                const uint32_t s0n = (first[e] >> 16) & 3u;
                const uint32_t rowmask = (tab.unsafe_rows >> (8 * s0n)) & 0xFFu;
                const uint32_t uns = ((first[e] >> 12) & 1u) | ((((orr[e] >> 8) & rowmask) != 0u) ? 1u : 0u);
                float rr = r[e];
                if (tab.reward_log2) rr = log2_1p(rr);
                rout[e] = rr;
                unsafe_w |= uns << (8 * e); count_w |= (add[e] & 31u) << (8 * e);
                if (e < rem) st_reward += __float2int_rn(rr * 16777216.0f);
            }
            const uint32_t vb = valid_bytes(rem);
            st_steps += rem;
            st_unsafe = add_bytes(unsafe_w & vb, st_unsafe);
            st_count = add_bytes(count_w & vb, st_count);
            st_trunc = add_bytes(trunc_w & vb, st_trunc);
            o_t[tid] = make_int4(tn[0], tn[1], tn[2], tn[3]);
            o_reward[tid] = make_int4(__float_as_int(rout[0]), __float_as_int(rout[1]), __float_as_int(rout[2]), __float_as_int(rout[3]));
            o_index[tid] = make_int4(idx[0], idx[1], idx[2], idx[3]);
            o_flags[tid] = 0u;
            o_flags[kThreads + tid] = trunc_w;
            o_flags[2 * kThreads + tid] = unsafe_w;
            o_flags[3 * kThreads + tid] = count_w;
        }
        fence_proxy_async();                                         // generic-proxy writes -> visible to the bulk stores
        __syncthreads();
        if (tid == 0) {
            const int64_t left = io.end - e_tile;
            const uint32_t nb = static_cast<uint32_t>(((left < kTile ? left : kTile) + 15) / 16 * 16);
            const uint32_t in_base = smem_u32(s_in + s * L::kInBytes), out_base = smem_u32(s_out);
#pragma unroll 1
            for (int c = 0; c < C; ++c) bulk_store(io.state + c * ld + e_tile, in_base + c * 1024, nb);
            bulk_store(io.t + e_tile, out_base, nb * 4);
            bulk_store(io.reward + e_tile, out_base + 4096, nb * 4);
            bulk_store(io.index + e_tile, out_base + 8192, nb * 4);
            bulk_store(io.terminated + e_tile, out_base + 12288, nb);
            bulk_store(io.truncated + e_tile, out_base + 12288 + 1024, nb);
            bulk_store(io.unsafe + e_tile, out_base + 12288 + 2048, nb);
            bulk_store(io.count + e_tile, out_base + 12288 + 3072, nb);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_all();
    if (io.stats) {
        const ThreadStats ts = {st_steps, st_unsafe, st_count, st_trunc, st_reward};
        block_flush_stats(ts, s_stats, io.stats);
    }
    tick_step_counter(io);
}

template <int C>
cudaError_t launch_tma_c(const CellTables &tab, const StepIO &io, const uint2 *lut, int n_sm, cudaStream_t st)
{
    using L = TmaLayout<C>;
    static int per_sm_dev[kMaxDevices] = {};     // per device: the shared-memory opt-in is a per-device attribute
    const int slot = current_device_slot();
    int per_sm = per_sm_dev[slot];
    if (per_sm == 0 || slot == kMaxDevices - 1) {
        cudaError_t e = cudaFuncSetAttribute(cell_tma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (e != cudaSuccess) return e;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cell_tma_kernel<C>, kThreads, L::kTotal) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        per_sm_dev[slot] = per_sm;
    }
    const int64_t tiles = (io.end - io.begin + kTile - 1) / kTile;
    const int64_t cap = static_cast<int64_t>(n_sm) * per_sm;
    const int grid = static_cast<int>(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
    cell_tma_kernel<C><<<grid, kThreads, L::kTotal, st>>>(tab, io, lut);
    return cudaGetLastError();
}

}  // namespace

// Deterministic, no side-effect rows, 8 <= C <= 16; the caller (gc_api.cu) checks.
cudaError_t gc_launch_cell_tma_step(const CellTables &tab, const StepIO &io, const uint2 *lut, int n_sm, cudaStream_t st)
{
    switch (tab.n_cells) {
#define GC_CASE(C) case C: return launch_tma_c<C>(tab, io, lut, n_sm, st);
        GC_CASE(8) GC_CASE(9) GC_CASE(10) GC_CASE(11) GC_CASE(12) GC_CASE(13) GC_CASE(14) GC_CASE(15) GC_CASE(16)
#undef GC_CASE
    default: return cudaErrorInvalidValue;
    }
}
