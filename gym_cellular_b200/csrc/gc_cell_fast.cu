// Fast path of the cellular (polarisation) step for S, A <= 4 levels, deterministic or stochastic.
//
// Two cells at a time: the four 2-bit digits (s_c, a_c, s_d, a_d) of a cell pair form one byte that,
// together with one "the noise draw of this cell fired" bit per cell in the stochastic case
// (cells3resetVdeadlock.py:35-61; Philox word < threshold, or a replayed uniform < p), indexes a table
// staged in shared memory (256 entries deterministic, 1024 stochastic; built on the host by
// gc_build_pair_lut).  One 64-bit shared load returns, for the pair,
//   .y  the reward contribution R[s_c][a_c] + R[s_d][a_d]                (cells3states3actions3.py:9-45)
//   .x  bits  0-4   how many of the two next levels count towards the incidence     (:159-162)
//       bits  8-11  one-hot set of the two next levels ("presence", for the side-effect report)
//       bit   12    row-0 entries 0 and 1 of the side-effects matrix hold 'unsafe' for (s'_c, s'_d)
//                   (meaningful for the pair (cell 0, cell 1) only)                 (:157-212)
//       bits 16-23  next level of cell c, bits 24-31 next level of cell d           (:133-154)
// so the per-(env, pair) work is: byte extract, LDS.64, FADD, IADD, LOP.  The 4 envs a thread owns
// travel through the pair-index arithmetic together (byte SWAR in one 32-bit word), and the next-state
// rows are rebuilt in SoA order with PRMT byte shuffles.  'unsafe' for the cells j >= 2 is evaluated
// per env from the presence set: unsafe iff some present level x has SE[j][s'_0][x] == unsafe, which
// needs the side-effect tables of all cells j >= 2 to be equal (they are for every table set
// gym_cellular_b200/tables.py builds; otherwise the generic kernel in gc_kernels.cu is used).
#include "gc_device.cuh"

namespace {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

// resident blocks per SM the register budget is sized for: all 2C row words of a thread's 4 envs are
// requested up front (memory-level parallelism), so wide envs trade occupancy for loads in flight
constexpr int pair_min_blocks(int c) { return c > 8 ? 2 : (c > 4 ? 3 : 4); }

template <int C, int RNG, bool WITH_SE>
__global__ void __launch_bounds__(kThreads, pair_min_blocks(C))
cell_pair_kernel(const __grid_constant__ CellTables tab, const __grid_constant__ StepIO io,
                 const uint2 *__restrict__ lut)
{
    constexpr int NP = C / 2;
    constexpr bool ODD = (C & 1) != 0;
    constexpr int N_PAIR = (RNG == GC_RNG_NONE) ? 256 : GC_PAIR_LUT_PAIRS;
    constexpr int N_SINGLE = (RNG == GC_RNG_NONE) ? 16 : 32;
    __shared__ uint2 s_pair[N_PAIR];
    __shared__ uint2 s_single[N_SINGLE];
    __shared__ uint8_t s_se[WITH_SE ? C : 1][GC_TBL];
    __shared__ unsigned long long s_stats[5];

    for (int i = threadIdx.x; i < N_PAIR; i += kThreads) s_pair[i] = lut[i];
    if (threadIdx.x < N_SINGLE) s_single[threadIdx.x] = lut[GC_PAIR_LUT_PAIRS + threadIdx.x];
    if (WITH_SE)
        for (int i = threadIdx.x; i < C * GC_TBL; i += kThreads) s_se[i / GC_TBL][i % GC_TBL] = tab.se[i / GC_TBL][i % GC_TBL];
    if (threadIdx.x < 5) s_stats[threadIdx.x] = 0;
    __syncthreads();

    ThreadStats ts = {0, 0, 0, 0, 0};
    const int64_t ld = io.ld;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads * kEPT;
    for (int64_t e0 = io.begin + (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * kEPT;
         e0 < io.end; e0 += stride) {
        uint32_t sw[C], aw[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            sw[c] = ld_stream_u32(io.state + c * ld + e0);
            aw[c] = ld_stream_u32(io.actions + c * ld + e0) & 0x03030303u;   // keep byte lanes apart
        }
        const int4 t4 = ld_stream_v4(io.t + e0);
        const int tin[kEPT] = {t4.x, t4.y, t4.z, t4.w};
        int tn[kEPT] = {t4.x + 1, t4.y + 1, t4.z + 1, t4.w + 1};
        uint32_t rnd[kEPT][4];                                     // Philox block of the current 4 cells
        uint32_t trunc_w = 0, keep = 0xFFFFFFFFu;                 // keep: byte mask of envs NOT reset
        if (io.max_episode_steps > 0) {
#pragma unroll
            for (int e = 0; e < kEPT; ++e)
                if (tn[e] >= io.max_episode_steps) { tn[e] = 0; trunc_w |= 1u << (8 * e); keep &= ~(0xFFu << (8 * e)); }
        }

        float r[kEPT] = {0.f, 0.f, 0.f, 0.f};
        uint32_t add[kEPT] = {0, 0, 0, 0}, orr[kEPT] = {0, 0, 0, 0}, first[kEPT];
        uint32_t idx[kEPT] = {0, 0, 0, 0};
        uint32_t raw[WITH_SE ? C : 1];
        uint32_t q = 0;                                            // 4 cells' worth of index digits per byte
#pragma unroll
        for (int k = 0; k < NP + (ODD ? 1 : 0); ++k) {
            const int c = 2 * k, d = 2 * k + 1;
            const bool pair = k < NP;
            uint32_t inf[kEPT];
            if (RNG == GC_RNG_PHILOX && (c & 3) == 0) {
#pragma unroll
                for (int e = 0; e < kEPT; ++e) {
                    const uint64_t gid = static_cast<uint64_t>(io.env_id_offset + e0 + e);
                    const uint32_t ctr = io.episodic ? static_cast<uint32_t>(tin[e]) : io.rng_counter;
                    philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), ctr,
                                  static_cast<uint32_t>(c >> 2), io.round_key, rnd[e]);
                }
            }
            // fire(e, cell): did the noise draw of that cell fire?  (ignored by the table where the
            // (level, action) pair consumes no draw)
            auto fire = [&](int e, int cell) -> uint32_t {
                if (RNG == GC_RNG_PHILOX) return (tab.noise_thr_nz && rnd[e][cell & 3] <= tab.noise_thr_m1) ? 1u : 0u;
                if (RNG == GC_RNG_REPLAY)
                    return ((e0 + e) < io.end && io.replay[(e0 + e) * C + cell] < tab.noise_prob) ? 1u : 0u;
                return 0u;
            };
            if (pair) {
                const uint32_t pidx = (aw[d] * 4u + (sw[d] & 0x03030303u)) * 16u + aw[c] * 4u + (sw[c] & 0x03030303u);
#pragma unroll
                for (int e = 0; e < kEPT; ++e) {
                    uint32_t ix = byte_of(pidx, e);
                    if (RNG != GC_RNG_NONE) ix |= (fire(e, c) << 8) | (fire(e, d) << 9);
                    const uint2 ent = s_pair[ix];
                    r[e] += __uint_as_float(ent.y);
                    inf[e] = ent.x;
                }
            } else {
                const uint32_t sidx = aw[c] * 4u + (sw[c] & 0x03030303u);
#pragma unroll
                for (int e = 0; e < kEPT; ++e) {
                    uint32_t ix = byte_of(sidx, e) & 15u;
                    if (RNG != GC_RNG_NONE) ix |= fire(e, c) << 4;
                    const uint2 ent = s_single[ix];
                    r[e] += __uint_as_float(ent.y);
                    inf[e] = ent.x;
                }
            }
#pragma unroll
            for (int e = 0; e < kEPT; ++e) {
                add[e] += inf[e];
                if (k == 0) first[e] = inf[e]; else orr[e] |= inf[e];
            }
            // SoA rows of the next state: byte 2 (cell c) and byte 3 (cell d) of the four info words
            const uint32_t u = prmt(inf[0], inf[1], 0x7362), v = prmt(inf[2], inf[3], 0x7362);
            const uint32_t row_c = prmt(u, v, 0x5410), row_d = prmt(u, v, 0x7632);
            if (WITH_SE) { raw[c] = row_c; if (pair) raw[d] = row_d; }
            const uint32_t out_c = (row_c & keep) | (0x01010101u * static_cast<uint8_t>(tab.init[c]) & ~keep);
            st_stream_u32(io.state + c * ld + e0, out_c);
            // index digits: q collects cells 4g..4g+3 as one base-S^4 digit per env byte
            q += out_c * tab.place4[c & 3];
            if (pair) {
                const uint32_t out_d = (row_d & keep) | (0x01010101u * static_cast<uint8_t>(tab.init[d]) & ~keep);
                st_stream_u32(io.state + d * ld + e0, out_d);
                q += out_d * tab.place4[d & 3];
            }
            if ((k & 1) == 1 || k == NP + (ODD ? 1 : 0) - 1) {     // a group of 4 cells is complete
#pragma unroll
                for (int e = 0; e < kEPT; ++e) idx[e] += byte_of(q, e) * tab.place[(c / 4) * 4];
                q = 0;
            }
        }

        uint32_t unsafe_w = 0, count_w = 0;
        float rout[kEPT];
#pragma unroll
        for (int e = 0; e < kEPT; ++e) {
            const uint32_t s0n = (first[e] >> 16) & 3u;
            const uint32_t rowmask = (tab.unsafe_rows >> (8 * s0n)) & 0xFFu;
            const uint32_t uns = ((first[e] >> 12) & 1u) | ((((orr[e] >> 8) & rowmask) != 0u) ? 1u : 0u);
            const uint32_t cnt = add[e] & 31u;
            float rr = r[e];
            if (tab.reward_log2) rr = log1pf(rr) * 1.44269504088896341f;
            rout[e] = rr;
            unsafe_w |= uns << (8 * e); count_w |= cnt << (8 * e);
            if ((e0 + e) < io.end) {
                ts.steps += 1; ts.unsafe += uns; ts.count += cnt; ts.truncated += (trunc_w >> (8 * e)) & 1u;
                ts.reward_q24 += __float2ll_rn(rr * 16777216.0f);
            }
        }
        if (WITH_SE) {
            // row 0 of the side-effects matrix from the (pre-reset) next state
#pragma unroll
            for (int c = 0; c < C; ++c) {
                uint32_t sew = 0;
#pragma unroll
                for (int e = 0; e < kEPT; ++e) {
                    const uint32_t s0n = byte_of(raw[0], e);
                    const uint32_t partner = byte_of(raw[c == 0 ? (C > 1 ? 1 : 0) : c], e);
                    sew |= static_cast<uint32_t>(s_se[c][(s0n * GC_LVL_PAD + partner) & (GC_TBL - 1)]) << (8 * e);
                }
                st_stream_u32(io.se_row + c * ld + e0, sew);
            }
        }
        st_stream_v4(io.t + e0, make_int4(tn[0], tn[1], tn[2], tn[3]));
        st_stream_v4(io.reward + e0, make_int4(__float_as_int(rout[0]), __float_as_int(rout[1]),
                                               __float_as_int(rout[2]), __float_as_int(rout[3])));
        st_stream_v4(io.index + e0, make_int4(idx[0], idx[1], idx[2], idx[3]));
        st_stream_u32(io.terminated + e0, 0u);
        st_stream_u32(io.truncated + e0, trunc_w);
        st_stream_u32(io.unsafe + e0, unsafe_w);
        st_stream_u32(io.count + e0, count_w);
    }
    if (io.stats) block_flush_stats(ts, s_stats, io.stats);
}

template <int C, int RNG>
cudaError_t launch_pair_cr(const CellTables &tab, const StepIO &io, const uint2 *lut, int n_sm, cudaStream_t st)
{
    const int64_t n = io.end - io.begin;
    if (io.se_row)
        cell_pair_kernel<C, RNG, true><<<grid_for<cell_pair_kernel<C, RNG, true>>(n, n_sm), kThreads, 0, st>>>(tab, io, lut);
    else
        cell_pair_kernel<C, RNG, false><<<grid_for<cell_pair_kernel<C, RNG, false>>(n, n_sm), kThreads, 0, st>>>(tab, io, lut);
    return cudaGetLastError();
}

template <int C>
cudaError_t launch_pair_c(const CellTables &tab, const StepIO &io, const uint2 *lut, int rng_mode, int n_sm,
                          cudaStream_t st)
{
    switch (rng_mode) {
    case GC_RNG_NONE: return launch_pair_cr<C, GC_RNG_NONE>(tab, io, lut, n_sm, st);
    case GC_RNG_PHILOX: return launch_pair_cr<C, GC_RNG_PHILOX>(tab, io, lut, n_sm, st);
    default: return launch_pair_cr<C, GC_RNG_REPLAY>(tab, io, lut, n_sm, st);
    }
}

}  // namespace

cudaError_t gc_launch_cell_pair_step(const CellTables &tab, const StepIO &io, const uint2 *lut, int rng_mode,
                                     int n_sm, cudaStream_t st)
{
    switch (tab.n_cells) {
#define GC_CASE(C) case C: return launch_pair_c<C>(tab, io, lut, rng_mode, n_sm, st);
        GC_CASE(1) GC_CASE(2) GC_CASE(3) GC_CASE(4) GC_CASE(5) GC_CASE(6) GC_CASE(7) GC_CASE(8)
        GC_CASE(9) GC_CASE(10) GC_CASE(11) GC_CASE(12) GC_CASE(13) GC_CASE(14) GC_CASE(15) GC_CASE(16)
#undef GC_CASE
    default: return cudaErrorInvalidValue;
    }
}
