// Cellular step for 5..8 levels / actions per cell (int8 layout): the pair-table idea of gc_cell_fast.cu with
// 3-bit digits.
//
// The (level, action) digits of a cell pair (s_c, a_c, s_d, a_d) form a 12-bit index into a 4096-entry table
// staged in shared memory (32 KB; built on the host by gc_build_pair8_lut from the same [S][A] tables,
// cells3states3actions3.py:9-45, 133-154, 157-212 generalised to S levels as in gym_cellular_b200/tables.py).
// One 64-bit shared load returns, for the pair,
//   .y  the reward contribution R[s_c][a_c] + R[s_d][a_d]
//   .x  bits  0-4   how many of the two next levels count towards the incidence
//       bit   7     row-0 entries 0 and 1 of the side-effects matrix hold 'unsafe' for (s'_c, s'_d)
//                   (meaningful for the pair (cell 0, cell 1) only)
//       bits  8-15  one-hot set of the two next levels (for the 'unsafe' report of the cells j >= 2)
//       bits 16-23  next level of cell c, bits 24-31 next level of cell d
// A 12-bit index does not fit the byte lanes the 2-bit kernel uses, so the four envs of a thread travel as two
// 16-bit-lane words (envs 0, 2 and envs 1, 3).  The generic per-cell kernel (gc_kernels.cu) did ~40 instructions
// per (env, cell) for these shapes and ran at 0.47-0.49 of the HBM roofline (10 cells x 8 levels).
//
// Stochastic envs (RNG == GC_RNG_PHILOX; cells3resetVdeadlock.py:35-61 generalised): two fire bits on top of a
// 12-bit pair index would need a 128 KB table, so the cells are looked up ONE at a time in a 128-entry table
// [fire][a][s] (same entry layout, cell d empty), and the 'unsafe' flag of the pair (cell 0, cell 1) comes from a
// 64-bit mask over (s'_0, s'_1) instead of the pair entry.  The draws are those of every other kernel: one 32-bit
// Philox word per cell up to GC_NARROW_CELLS cells, 16-bit halves with the exact tie rule beyond (fire_bits_wide).
// The generic kernel ran these shapes at 0.22 of the HBM roofline (10 cells x 8 levels with noise, 570 us per step
// of 2^24 envs).
// WITH_SE: row 0 of the side-effects matrix is written as well (entry j from (s'_0, s'_p), p = 1 for j = 0 and
// p = j otherwise; cells3states3actions3.py:157-212), one byte lookup per (env, cell) in the per-cell tables.
// Replayed draws and ragged radices stay with the generic kernel.
#include "gc_device.cuh"

namespace {

#ifndef GC_PAIR8_MINB
#define GC_PAIR8_MINB 4
#endif

template <int RNG, bool WITH_SE>
#ifndef GC_PAIR8_NOISE_MINB
#define GC_PAIR8_NOISE_MINB 4        // 64 registers (16 bytes spilled): 311 us against 317 us at three blocks, 10 x 8 with noise
#endif
__global__ void __launch_bounds__(kThreads, RNG == GC_RNG_NONE ? GC_PAIR8_MINB : GC_PAIR8_NOISE_MINB)
cell_pair8_kernel(const __grid_constant__ CellTables tab, const __grid_constant__ StepIO io, const uint2 *__restrict__ lut)
{
    constexpr bool NOISE = RNG == GC_RNG_PHILOX;
    constexpr int N_SINGLE = NOISE ? 128 : 64;
    __shared__ uint2 s_pair[NOISE ? 1 : GC_PAIR8_PAIRS];
    __shared__ uint2 s_single[N_SINGLE];
    __shared__ uint8_t s_se[WITH_SE ? GC_MAX_CELLS : 1][GC_TBL];
    __shared__ unsigned long long s_stats[5];
    __shared__ StepCounterShared s_ctr;
    const int C = tab.n_cells;
    // 32-bit element indexes (n_cells * ld <= 2^31, gc_create): an address is one IMAD.WIDE.U32 on the FMA pipe
    const uint32_t ld = static_cast<uint32_t>(io.ld);
    const uint32_t stride = gridDim.x * kThreads * kEPT, e_end = static_cast<uint32_t>(io.end);
    uint32_t e0 = static_cast<uint32_t>(io.begin) + (blockIdx.x * kThreads + threadIdx.x) * kEPT;

    if constexpr (!NOISE)
        for (int i = threadIdx.x; i < GC_PAIR8_PAIRS; i += kThreads) s_pair[i] = lut[i];
    if (threadIdx.x < N_SINGLE) s_single[threadIdx.x] = lut[GC_PAIR8_PAIRS + threadIdx.x];
    if constexpr (WITH_SE)
        for (int i = threadIdx.x; i < tab.n_cells * GC_TBL; i += kThreads) s_se[i / GC_TBL][i % GC_TBL] = tab.se[i / GC_TBL][i % GC_TBL];
    if (threadIdx.x < 5) s_stats[threadIdx.x] = 0;
    pdl_launch_dependents();
    pdl_wait();
    step_counter_read(io, &s_ctr);
    __syncthreads();
    [[maybe_unused]] const uint32_t step_counter = step_counter_arrive(io, &s_ctr);

    uint32_t st_steps = 0, st_unsafe = 0, st_count = 0, st_trunc = 0;
    long long st_reward = 0;
#pragma unroll 1
    for (; e0 < e_end; e0 += stride) {
        const int rem = static_cast<int>(e_end - e0 < kEPT ? e_end - e0 : kEPT);
        // rows of the first four cells, the episode steps
        uint32_t sw[4] = {0, 0, 0, 0}, aw[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < C) { sw[i] = ld_stream_u32(io.state + (i * ld + e0)); aw[i] = ld_stream_u32(io.actions + (i * ld + e0)); }
        const int4 t4 = ld_stream_v4(io.t + e0);
        int tn[kEPT] = {t4.x + 1, t4.y + 1, t4.z + 1, t4.w + 1};
        uint32_t trunc_w = 0, keep = 0xFFFFFFFFu;
        if (io.max_episode_steps > 0) {
#pragma unroll
            for (int e = 0; e < kEPT; ++e)
                if (tn[e] >= io.max_episode_steps) { tn[e] = 0; trunc_w |= 1u << (8 * e); keep &= ~(0xFFu << (8 * e)); }
        }
        float r[kEPT] = {0.f, 0.f, 0.f, 0.f};
        uint32_t idx[kEPT] = {0, 0, 0, 0};
        uint32_t sum01 = 0, sum23 = 0;          // info low halves of envs (0, 1) and (2, 3) in 16-bit lanes: count in bits 0-4
        uint32_t or01 = 0, or23 = 0;            // bits 8-15 of a lane: levels present in the cells j >= 2; bit 7: pair (0, 1) flag
        uint32_t s0w = 0;                       // next level of cell 0, byte lane e = env e
        [[maybe_unused]] uint32_t s1w = 0;      // next level of cell 1 (stochastic variant)
        // fire[e]: bit c = the noise draw of cell c of env e fired (the table ignores it where (level, action)
        // consumes no draw)
        [[maybe_unused]] uint32_t fire[kEPT] = {0, 0, 0, 0};
        if constexpr (NOISE) {
            const uint64_t gid0 = static_cast<uint64_t>(io.env_id_offset) + e0;
            const uint32_t gid_lo = static_cast<uint32_t>(gid0), gid_hi = static_cast<uint32_t>(gid0 >> 32);
            const int tin[kEPT] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int e = 0; e < kEPT; ++e) {
                const uint32_t ctr = io.episodic ? static_cast<uint32_t>(tin[e]) : step_counter;
                if (C > GC_NARROW_CELLS) {
                    fire[e] = fire_bits_wide<2>(tab, gid_lo | e, gid_hi, ctr, io.round_key);
                } else {
                    uint32_t w[4];
                    philox4x32_10(gid_lo | e, gid_hi, ctr, 0u, io.round_key, w);
                    const uint32_t thr = tab.noise_thr_m1;    // threshold - 1; a zero threshold never gets here (gc_api.cu)
                    fire[e] = (w[0] <= thr ? 1u : 0u) | (w[1] <= thr ? 2u : 0u) | (w[2] <= thr ? 4u : 0u) | (w[3] <= thr ? 8u : 0u);
                }
            }
        }

        // entry j of row 0 of the side-effects matrix, from (s'_0, s'_p) before an auto-reset
        [[maybe_unused]] auto emit_se = [&](int j, uint32_t partner_row) {
            uint32_t sew = 0;
#pragma unroll
            for (int e = 0; e < kEPT; ++e)
                sew |= static_cast<uint32_t>(s_se[j][((byte_of(s0w, e) & 7u) * GC_LVL_PAD + (byte_of(partner_row, e) & 7u))]) << (8 * e);
            st_stream_u32(io.se_row + (j * ld + e0), sew);
        };
        // bookkeeping shared by a pair and a single cell: info low halves, next-state rows of cell c (and d)
        auto account = [&](const uint2 (&ent)[kEPT], int c, bool pair, uint32_t keep_bits) {
#pragma unroll
            for (int e = 0; e < kEPT; ++e) r[e] += __uint_as_float(ent[e].y);       // cell order, from 0.0
            const uint32_t w01 = prmt(ent[0].x, ent[1].x, 0x5410), w23 = prmt(ent[2].x, ent[3].x, 0x5410);
            sum01 += w01 & 0x001F001Fu; sum23 += w23 & 0x001F001Fu;
            or01 |= w01 & keep_bits; or23 |= w23 & keep_bits;
            // SoA rows of the next state: byte 2 (cell c) and byte 3 (cell d) of the four info words
            const uint32_t u = prmt(ent[0].x, ent[1].x, 0x7362), v = prmt(ent[2].x, ent[3].x, 0x7362);
            const uint32_t row_c = prmt(u, v, 0x5410), row_d = prmt(u, v, 0x7632);
            if (c == 0) s0w = row_c;
            if (NOISE && c == 1) s1w = row_c;
            if constexpr (WITH_SE) {
                if (c == 0) {
                    if (pair) { emit_se(0, row_d); emit_se(1, row_d); }
                    else if (C == 1) emit_se(0, row_c);
                } else if (c == 1) {                       // single-cell lookups only (pairs start at even cells)
                    emit_se(0, row_c); emit_se(1, row_c);
                } else {
                    emit_se(c, row_c);
                    if (pair) emit_se(c + 1, row_d);
                }
            }
            const uint32_t out_c = (row_c & keep) | ((0x01010101u * static_cast<uint8_t>(tab.init[c])) & ~keep);
            st_stream_u32(io.state + (c * ld + e0), out_c);
            if (io.final_state) st_stream_u32(io.final_state + (c * ld + e0), row_c);
            const uint32_t pc = tab.place[c];
#pragma unroll
            for (int e = 0; e < kEPT; ++e) idx[e] += byte_of(out_c, e) * pc;
            if (pair) {
                const uint32_t out_d = (row_d & keep) | ((0x01010101u * static_cast<uint8_t>(tab.init[c + 1])) & ~keep);
                st_stream_u32(io.state + ((c + 1) * ld + e0), out_d);
                if (io.final_state) st_stream_u32(io.final_state + ((c + 1) * ld + e0), row_d);
                const uint32_t pd = tab.place[c + 1];
#pragma unroll
                for (int e = 0; e < kEPT; ++e) idx[e] += byte_of(out_d, e) * pd;
            }
        };
        // cells c, c + 1 (rows i, i + 1 of the current group): a pair lookup, or a single-cell lookup for a last odd cell
        auto do_pair = [&](int c, int i) {
            const uint32_t code_c = (aw[i] & 0x07070707u) * 8u + (sw[i] & 0x07070707u);     // 6-bit (level, action) code per env
            uint2 ent[kEPT];
            if (c + 1 < C) {
                const uint32_t code_d = (aw[i + 1] & 0x07070707u) * 8u + (sw[i + 1] & 0x07070707u);
                // the 12-bit pair index in 16-bit lanes: envs (0, 2) and envs (1, 3)
                const uint32_t p02 = (code_c & 0x00FF00FFu) + (code_d & 0x00FF00FFu) * 64u;
                const uint32_t p13 = ((code_c >> 8) & 0x00FF00FFu) + ((code_d >> 8) & 0x00FF00FFu) * 64u;
                ent[0] = s_pair[p02 & 0xFFFu]; ent[2] = s_pair[(p02 >> 16) & 0xFFFu];
                ent[1] = s_pair[p13 & 0xFFFu]; ent[3] = s_pair[(p13 >> 16) & 0xFFFu];
                account(ent, c, true, c == 0 ? 0x00800080u : 0xFF00FF00u);
            } else {
#pragma unroll
                for (int e = 0; e < kEPT; ++e) ent[e] = s_single[byte_of(code_c, e) & 63u];
                account(ent, c, false, c == 0 ? 0x00800080u : 0xFF00FF00u);
            }
        };
        // stochastic variant: cell c (row i of the current group) alone, its fire bit on top of the 6-bit code
        auto do_single = [&](int c, int i) {
            const uint32_t code = (aw[i] & 0x07070707u) * 8u + (sw[i] & 0x07070707u);
            uint2 ent[kEPT];
#pragma unroll
            for (int e = 0; e < kEPT; ++e) ent[e] = s_single[(byte_of(code, e) & 63u) | (((fire[e] >> c) & 1u) << 6)];
            account(ent, c, false, c >= 2 ? 0xFF00FF00u : 0u);
        };

        // four cells (two lookups per env) per iteration; the rows of the next four are requested before these are
        // computed, so that a thread keeps up to 16 row words in flight
#pragma unroll 1
        for (int c = 0; c < C; c += 4) {
            uint32_t ns[4] = {0, 0, 0, 0}, na[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (c + 4 + i < C) {
                    ns[i] = ld_stream_u32(io.state + ((c + 4 + i) * ld + e0));
                    na[i] = ld_stream_u32(io.actions + ((c + 4 + i) * ld + e0));
                }
            if constexpr (NOISE) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (c + i < C) do_single(c + i, i);
            } else {
                do_pair(c, 0);
                if (c + 2 < C) do_pair(c + 2, 2);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { sw[i] = ns[i]; aw[i] = na[i]; }
        }

        // unsafe / count per env: the unsafe-levels mask of s'_0 against the levels present in the cells j >= 2
        const uint32_t count_w = prmt(sum01, sum23, 0x6420) & 0x1F1F1F1Fu;
        const uint32_t present = prmt(or01, or23, 0x7531);                 // byte e: levels present in cells j >= 2 of env e
        uint32_t unsafe_w = NOISE ? 0u : (prmt(or01, or23, 0x6420) >> 7) & 0x01010101u;
#pragma unroll
        for (int e = 0; e < kEPT; ++e) {
            if constexpr (NOISE) {      // entries 0 and 1 of row 0 of the side-effects matrix, from (s'_0, s'_1) (s'_0 twice if C == 1)
                const uint32_t s1n = byte_of(C >= 2 ? s1w : s0w, e) & 7u;
                unsafe_w |= (static_cast<uint32_t>(tab.unsafe01_rows8 >> (8u * (byte_of(s0w, e) & 7u) + s1n)) & 1u) << (8 * e);
            }
            const uint32_t rowmask = static_cast<uint32_t>(tab.unsafe_rows8 >> (8u * (byte_of(s0w, e) & 7u))) & 0xFFu;
            unsafe_w |= ((byte_of(present, e) & rowmask) ? 1u : 0u) << (8 * e);
        }
        float rout[kEPT];
        log2_1p_x4(tab.reward_log2, r, rout);
#pragma unroll
        for (int e = 0; e < kEPT; ++e)
            if (e < rem) st_reward += __float2int_rn(rout[e] * 16777216.0f);
        {
            const uint32_t vb = valid_bytes(rem);
            st_steps += rem;
            st_unsafe = add_bytes(unsafe_w & vb, st_unsafe);
            st_count = add_bytes(count_w & vb, st_count);
            st_trunc = add_bytes(trunc_w & vb, st_trunc);
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
    if (io.stats) {
        const ThreadStats ts = {st_steps, st_unsafe, st_count, st_trunc, st_reward};
        block_flush_stats(ts, s_stats, io.stats);
    }
    step_counter_finish(io, &s_ctr);
}

}  // namespace

cudaError_t gc_launch_cell_pair8_step(const CellTables &tab, const StepIO &io, const uint2 *lut, int rng_mode, int n_sm,
                                      cudaStream_t st)
{
    const int64_t n = io.end - io.begin;
#define GC_PAIR8_LAUNCH(RNG, SE) \
    return launch_step_kernel(cell_pair8_kernel<RNG, SE>, grid_for<cell_pair8_kernel<RNG, SE>>(n, n_sm), kThreads, 0, st, tab, io, lut)
    if (rng_mode == GC_RNG_PHILOX) {
        if (io.se_row) GC_PAIR8_LAUNCH(GC_RNG_PHILOX, true);
        GC_PAIR8_LAUNCH(GC_RNG_PHILOX, false);
    }
    if (io.se_row) GC_PAIR8_LAUNCH(GC_RNG_NONE, true);
    GC_PAIR8_LAUNCH(GC_RNG_NONE, false);
#undef GC_PAIR8_LAUNCH
}
