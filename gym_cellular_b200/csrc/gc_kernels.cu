// sm_100a kernels of the batched gym-cellular step.
//
// Layout and mapping.  Every int8 row (one cell of the state / action, [ld] envs) is moved as
// 32-bit words: a thread owns kEPT = 4 consecutive envs, so a warp reads 128 contiguous bytes of each
// int8 row and 512 contiguous bytes (one 128-bit access per lane) of each 32-bit per-env vector
// (t, reward, index).  All accesses are fully coalesced, every byte is touched exactly once, and the
// block loops grid-stride over the env range with a grid of (SM count x resident blocks).
//
// The work is HBM-bound integer/byte arithmetic (3C+20 bytes per env-step); there is
// no contraction, hence no tensor-core path.  Transition / reward / side-effect tables arrive as a
// __grid_constant__ parameter block and are staged into shared memory once per block.
//
// Reference behaviour implemented here (paths relative to the reference checkout):
//   cellular step   gym_cellular/envs/cells3states3actions3.py:116-212, cells2rest3.py:103-187,
//                   cells3resetVdeadlock.py:35-68,148-228 (via the tables built in
//                   gym_cellular_b200/tables.py)
//   (grid world: gc_grid.cu; fast cellular path for S, A <= 4: gc_cell_fast.cu)
//   codec           gym_cellular/envs/utils/generalized_space_transformations.py:1-23
#include "gc_device.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// Generic cellular step: any table set with up to GC_MAX_LEVELS levels / actions and per-cell side-effect
// tables (the fast pair-table kernel of gc_cell_fast.cu covers n_states, n_actions <= 4 with equal tables
// for the cells j >= 2).  One rolled loop over the cells; the row words of cell c+1 are requested before
// cell c is computed; per (env, cell) one 64-bit shared load for the (level, action) entry, one for the
// side-effect code.  ~50 registers, so eight blocks of 256 threads stay resident per SM.
template <int RNG>
__global__ void __launch_bounds__(kThreads, RNG == GC_RNG_PHILOX ? 3 : 4)      // Philox variant: 85 registers, no spills
cell_step_kernel(const __grid_constant__ CellTables tab, const __grid_constant__ StepIO io)
{
    __shared__ uint2 s_sa[GC_TBL];                 // .x packed move/noisy/draws, .y reward bits
    __shared__ float s_rn[GC_TBL];                 // reward when the draw fired
    __shared__ uint8_t s_se[GC_MAX_CELLS][GC_TBL];
    __shared__ unsigned long long s_stats[5];
    const int C = tab.n_cells;

    for (int i = threadIdx.x; i < GC_TBL; i += kThreads)
        s_sa[i] = make_uint2(tab.sa[i], __float_as_uint(tab.reward[i])), s_rn[i] = tab.reward_noisy[i];
    for (int i = threadIdx.x; i < C * GC_TBL; i += kThreads)
        s_se[i / GC_TBL][i % GC_TBL] = tab.se[i / GC_TBL][i % GC_TBL];
    if (threadIdx.x < 5) s_stats[threadIdx.x] = 0;
    __syncthreads();

    const uint32_t step_counter = (RNG == GC_RNG_PHILOX) ? launch_step_counter(io) : 0u;
    uint32_t st_steps = 0, st_unsafe = 0, st_count = 0, st_trunc = 0;
    long long st_reward = 0;
    // 32-bit element indexes (n_cells * ld <= 2^31, gc_create): an address is one IMAD.WIDE.U32 on the FMA pipe
    const uint32_t ld = static_cast<uint32_t>(io.ld);
    const uint32_t stride = gridDim.x * kThreads * kEPT, e_end = static_cast<uint32_t>(io.end);
    for (uint32_t e0 = static_cast<uint32_t>(io.begin) + (blockIdx.x * kThreads + threadIdx.x) * kEPT;
         e0 < e_end; e0 += stride) {
        const int rem = static_cast<int>(e_end - e0 < kEPT ? e_end - e0 : kEPT);
        const uint64_t gid0 = static_cast<uint64_t>(io.env_id_offset) + e0;
        const uint32_t gid_lo = static_cast<uint32_t>(gid0), gid_hi = static_cast<uint32_t>(gid0 >> 32);
        uint32_t sw = ld_stream_u32(io.state + e0), aw = ld_stream_u32(io.actions + e0);
        const int4 t4 = ld_stream_v4(io.t + e0);
        const int tin[kEPT] = {t4.x, t4.y, t4.z, t4.w};
        int tn[kEPT] = {t4.x + 1, t4.y + 1, t4.z + 1, t4.w + 1};
        uint32_t trunc_w = 0, keep = 0xFFFFFFFFu;
        if (io.max_episode_steps > 0) {
#pragma unroll
            for (int e = 0; e < kEPT; ++e)
                if (tn[e] >= io.max_episode_steps) { tn[e] = 0; trunc_w |= 1u << (8 * e); keep &= ~(0xFFu << (8 * e)); }
        }
        float r[kEPT] = {0.f, 0.f, 0.f, 0.f};
        uint32_t idx[kEPT] = {0, 0, 0, 0};
        uint32_t unsafe_w = 0, count_w = 0, row0 = 0;        // row0: next levels of cell 0 (byte lanes)
        uint32_t rnd[kEPT][4];
        uint32_t fire16[kEPT] = {0, 0, 0, 0};
        const bool wide = C > GC_NARROW_CELLS;
        if (RNG == GC_RNG_PHILOX && wide) {                   // wide env: one Philox block per env (fire_bits_wide)
#pragma unroll
            for (int e = 0; e < kEPT; ++e)
                fire16[e] = fire_bits_wide<2>(tab, gid_lo | e, gid_hi, io.episodic ? static_cast<uint32_t>(tin[e]) : step_counter,
                                              io.round_key);
        }

#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            uint32_t sn = 0, an = 0;
            if (c + 1 < C) {                                  // prefetch the next cell's rows
                sn = ld_stream_u32(io.state + ((c + 1) * ld + e0));
                an = ld_stream_u32(io.actions + ((c + 1) * ld + e0));
            }
            if (RNG == GC_RNG_PHILOX && !wide && (c & 3) == 0) {
#pragma unroll
                for (int e = 0; e < kEPT; ++e) {
                    const uint32_t ctr = io.episodic ? static_cast<uint32_t>(tin[e]) : step_counter;
                    philox4x32_10(gid_lo | e, gid_hi, ctr, static_cast<uint32_t>(c >> 2), io.round_key, rnd[e]);
                }
            }
            uint32_t row = 0, sew = 0;
#pragma unroll
            for (int e = 0; e < kEPT; ++e) {
                const uint32_t sa_ix = (byte_of(sw, e) * GC_LVL_PAD + byte_of(aw, e)) & (GC_TBL - 1);
                const uint2 ent = s_sa[sa_ix];
                bool fire = false;
                if (RNG == GC_RNG_PHILOX) {
                    const uint32_t word = (c & 3) == 0 ? rnd[e][0] : (c & 3) == 1 ? rnd[e][1] : (c & 3) == 2 ? rnd[e][2] : rnd[e][3];
                    fire = (ent.x & 0x100u) && tab.noise_thr_nz && (wide ? ((fire16[e] >> c) & 1u) != 0u : word <= tab.noise_thr_m1);
                } else if (RNG == GC_RNG_REPLAY) {
                    if (e < rem && (ent.x & 0x100u)) fire = io.replay[static_cast<size_t>(e0 + e) * C + c] < tab.noise_prob;
                }
                const uint32_t nxt = fire ? ((ent.x >> 4) & 15u) : (ent.x & 15u);
                r[e] += fire ? s_rn[sa_ix] : __uint_as_float(ent.y);          // cell order, from 0.0
                row |= nxt << (8 * e);
                count_w += ((tab.counted_mask >> nxt) & 1u) << (8 * e);
            }
            if (c == 0) row0 = row;
            // row 0 of the side-effects matrix: entry j from (s'_0, s'_p), p = 1 for j = 0 and p = j
            // otherwise, so entry 0 is evaluated together with cell 1 (or with cell 0 itself if C == 1)
            uint32_t sew0 = 0;
#pragma unroll
            for (int e = 0; e < kEPT; ++e) {
                const uint32_t s0n = byte_of(row0, e), sp = byte_of(row, e);
                if (c > 0) {
                    const uint32_t code = s_se[c][(s0n * GC_LVL_PAD + sp) & (GC_TBL - 1)];
                    sew |= code << (8 * e);
                    unsafe_w |= (code == 2u ? 1u : 0u) << (8 * e);
                }
                if (c == 1 || C == 1) {
                    const uint32_t code0 = s_se[0][(s0n * GC_LVL_PAD + sp) & (GC_TBL - 1)];
                    sew0 |= code0 << (8 * e);
                    unsafe_w |= (code0 == 2u ? 1u : 0u) << (8 * e);
                }
            }
            if (io.se_row) {
                if (c > 0) st_stream_u32(io.se_row + (c * ld + e0), sew);
                if (c == 1 || C == 1) st_stream_u32(io.se_row + e0, sew0);
            }
            const uint32_t out = (row & keep) | ((0x01010101u * static_cast<uint8_t>(tab.init[c])) & ~keep);
            st_stream_u32(io.state + (c * ld + e0), out);
            if (io.final_state) st_stream_u32(io.final_state + (c * ld + e0), row);
            const uint32_t place = tab.place[c];
#pragma unroll
            for (int e = 0; e < kEPT; ++e) idx[e] += byte_of(out, e) * place;
            sw = sn; aw = an;
        }

        float rout[kEPT];
#pragma unroll
        for (int e = 0; e < kEPT; ++e) {
            float rr = r[e];
            if (tab.reward_log2) rr = log2_1p(rr);
            rout[e] = rr;
            if (e < rem) st_reward += __float2int_rn(rr * 16777216.0f);
        }
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
    tick_step_counter(io);
}

// ---------------------------------------------------------------------------------------------
// reset(): initial state, t = 0, tabular index (cells3states3actions3.py:99-113, grid_world.py:97-104)
struct InitBlock { int8_t cells[GC_MAX_CELLS]; uint32_t index; int32_t n_cells; };

__global__ void __launch_bounds__(kThreads)
reset_kernel(const __grid_constant__ InitBlock ini, const uint8_t *__restrict__ mask, int8_t *state,
             int32_t *t, uint32_t *index, int64_t n, int64_t ld)
{
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads * kEPT;
    for (int64_t e0 = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * kEPT; e0 < n; e0 += stride) {
        if (mask == nullptr) {
            for (int c = 0; c < ini.n_cells; ++c)
                *reinterpret_cast<uint32_t *>(state + c * ld + e0) = 0x01010101u * static_cast<uint8_t>(ini.cells[c]);
            *reinterpret_cast<int4 *>(t + e0) = make_int4(0, 0, 0, 0);
            if (index) *reinterpret_cast<uint4 *>(index + e0) = make_uint4(ini.index, ini.index, ini.index, ini.index);
        } else {
            const uint32_t m = *reinterpret_cast<const uint32_t *>(mask + e0);
            if (m == 0u) continue;
            // per-byte select mask: 0xFF where the mask byte is non-zero
            uint32_t sel = 0;
#pragma unroll
            for (int e = 0; e < kEPT; ++e) if (byte_of(m, e)) sel |= 0xFFu << (8 * e);
            for (int c = 0; c < ini.n_cells; ++c) {
                uint32_t *w = reinterpret_cast<uint32_t *>(state + c * ld + e0);
                *w = (*w & ~sel) | ((0x01010101u * static_cast<uint8_t>(ini.cells[c])) & sel);
            }
#pragma unroll
            for (int e = 0; e < kEPT; ++e)
                if (byte_of(m, e)) { t[e0 + e] = 0; if (index) index[e0 + e] = ini.index; }
        }
    }
}

// advances the device-resident global step (after the last chunk of a chunked pass)
__global__ void tick_kernel(uint32_t *step, uint32_t inc) { *step += inc; }

// Batched mixed-radix codec, uniform radix (generalized_space_transformations.py:1-23).
// index = sum_c cells[c] * radix^c  (cell 0 least significant); 32-bit unsigned arithmetic.
__global__ void __launch_bounds__(kThreads)
encode_kernel(int64_t n, int64_t ld, int n_cells, uint32_t radix, const int8_t *__restrict__ cells,
              uint32_t *__restrict__ index)
{
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads * kEPT;
    for (int64_t e0 = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * kEPT; e0 < n; e0 += stride) {
        uint32_t idx[kEPT] = {0, 0, 0, 0};
        uint32_t place = 1;
        for (int c = 0; c < n_cells; ++c) {
            const uint32_t w = ld_stream_u32(cells + c * ld + e0);
#pragma unroll
            for (int e = 0; e < kEPT; ++e) idx[e] += byte_of(w, e) * place;
            place *= radix;
        }
        st_stream_v4(index + e0, make_int4(idx[0], idx[1], idx[2], idx[3]));
    }
}

__global__ void __launch_bounds__(kThreads)
decode_kernel(int64_t n, int64_t ld, int n_cells, uint32_t radix, const uint32_t *__restrict__ index,
              int8_t *__restrict__ cells)
{
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads * kEPT;
    for (int64_t e0 = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * kEPT; e0 < n; e0 += stride) {
        const int4 v = ld_stream_v4(index + e0);
        uint32_t idx[kEPT] = {static_cast<uint32_t>(v.x), static_cast<uint32_t>(v.y),
                              static_cast<uint32_t>(v.z), static_cast<uint32_t>(v.w)};
        for (int c = 0; c < n_cells; ++c) {
            uint32_t w = 0;
#pragma unroll
            for (int e = 0; e < kEPT; ++e) {
                w |= (idx[e] % radix) << (8 * e);
                idx[e] /= radix;
            }
            st_stream_u32(cells + c * ld + e0, w);
        }
    }
}

// The reference's codec proper: a per-cell space list with arbitrary minimum and length
// (generalized_space_transformations.py:1-12: digit c = cells[c] - min(space[c]), radix c = len(space[c]);
// :15-23: the inverse by repeated modulo / floor division), batched.  32-bit unsigned index.
struct MixedRadix { uint32_t radix[GC_MAX_CELLS]; int32_t min[GC_MAX_CELLS]; int32_t n_cells; };

__global__ void __launch_bounds__(kThreads)
encode_mixed_kernel(const __grid_constant__ MixedRadix mr, int64_t n, int64_t ld, const int8_t *__restrict__ cells,
                    uint32_t *__restrict__ index)
{
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads * kEPT;
    for (int64_t e0 = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * kEPT; e0 < n; e0 += stride) {
        uint32_t idx[kEPT] = {0, 0, 0, 0};
        uint32_t place = 1;
        for (int c = 0; c < mr.n_cells; ++c) {
            const uint32_t w = ld_stream_u32(cells + c * ld + e0);
#pragma unroll
            for (int e = 0; e < kEPT; ++e) {
                const int digit = static_cast<int>(static_cast<int8_t>(byte_of(w, e))) - mr.min[c];
                idx[e] += static_cast<uint32_t>(digit) * place;
            }
            place *= mr.radix[c];
        }
        st_stream_v4(index + e0, make_int4(idx[0], idx[1], idx[2], idx[3]));
    }
}

__global__ void __launch_bounds__(kThreads)
decode_mixed_kernel(const __grid_constant__ MixedRadix mr, int64_t n, int64_t ld, const uint32_t *__restrict__ index,
                    int8_t *__restrict__ cells)
{
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads * kEPT;
    for (int64_t e0 = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * kEPT; e0 < n; e0 += stride) {
        const int4 v = ld_stream_v4(index + e0);
        uint32_t idx[kEPT] = {static_cast<uint32_t>(v.x), static_cast<uint32_t>(v.y),
                              static_cast<uint32_t>(v.z), static_cast<uint32_t>(v.w)};
        for (int c = 0; c < mr.n_cells; ++c) {
            const uint32_t radix = mr.radix[c];
            uint32_t w = 0;
#pragma unroll
            for (int e = 0; e < kEPT; ++e) {
                const int level = static_cast<int>(idx[e] % radix) + mr.min[c];
                w |= (static_cast<uint32_t>(level) & 0xFFu) << (8 * e);
                idx[e] /= radix;
            }
            st_stream_u32(cells + c * ld + e0, w);
        }
    }
}

}  // namespace

cudaError_t gc_launch_encode_mixed(int64_t n, int64_t ld, int n_cells, const int32_t *radix, const int32_t *min,
                                   const int8_t *cells, uint32_t *index, cudaStream_t st)
{
    MixedRadix mr = {};
    for (int c = 0; c < n_cells; ++c) { mr.radix[c] = static_cast<uint32_t>(radix[c]); mr.min[c] = min ? min[c] : 0; }
    mr.n_cells = n_cells;
    const int64_t need = (n + kThreads * kEPT - 1) / (kThreads * kEPT);
    encode_mixed_kernel<<<static_cast<int>(need < 1 ? 1 : (need > 65535 ? 65535 : need)), kThreads, 0, st>>>(mr, n, ld, cells, index);
    return cudaGetLastError();
}

cudaError_t gc_launch_decode_mixed(int64_t n, int64_t ld, int n_cells, const int32_t *radix, const int32_t *min,
                                   const uint32_t *index, int8_t *cells, cudaStream_t st)
{
    MixedRadix mr = {};
    for (int c = 0; c < n_cells; ++c) { mr.radix[c] = static_cast<uint32_t>(radix[c]); mr.min[c] = min ? min[c] : 0; }
    mr.n_cells = n_cells;
    const int64_t need = (n + kThreads * kEPT - 1) / (kThreads * kEPT);
    decode_mixed_kernel<<<static_cast<int>(need < 1 ? 1 : (need > 65535 ? 65535 : need)), kThreads, 0, st>>>(mr, n, ld, index, cells);
    return cudaGetLastError();
}

cudaError_t gc_launch_cell_step(const CellTables &tab, const StepIO &io, int rng_mode, int n_sm, cudaStream_t st)
{
    if (tab.n_cells < 1 || tab.n_cells > GC_MAX_CELLS) return cudaErrorInvalidValue;
    const int64_t n = io.end - io.begin;
    switch (rng_mode) {
    case GC_RNG_NONE:
        cell_step_kernel<GC_RNG_NONE><<<grid_for<cell_step_kernel<GC_RNG_NONE>>(n, n_sm), kThreads, 0, st>>>(tab, io);
        break;
    case GC_RNG_PHILOX:
        cell_step_kernel<GC_RNG_PHILOX><<<grid_for<cell_step_kernel<GC_RNG_PHILOX>>(n, n_sm), kThreads, 0, st>>>(tab, io);
        break;
    default:
        cell_step_kernel<GC_RNG_REPLAY><<<grid_for<cell_step_kernel<GC_RNG_REPLAY>>(n, n_sm), kThreads, 0, st>>>(tab, io);
        break;
    }
    return cudaGetLastError();
}

cudaError_t gc_launch_reset(int n_cells, const int8_t *init, uint32_t init_index, const uint8_t *mask,
                            int8_t *state, int32_t *t, uint32_t *index, int64_t n, int64_t ld, cudaStream_t st)
{
    InitBlock ini;
    for (int c = 0; c < GC_MAX_CELLS; ++c) ini.cells[c] = c < n_cells ? init[c] : 0;
    ini.index = init_index;
    ini.n_cells = n_cells;
    const int64_t need = (n + kThreads * kEPT - 1) / (kThreads * kEPT);
    reset_kernel<<<static_cast<int>(need < 1 ? 1 : (need > 65535 ? 65535 : need)), kThreads, 0, st>>>(
        ini, mask, state, t, index, n, ld);
    return cudaGetLastError();
}

cudaError_t gc_launch_tick(uint32_t *d_step, uint32_t inc, cudaStream_t st)
{
    tick_kernel<<<1, 1, 0, st>>>(d_step, inc);
    return cudaGetLastError();
}

cudaError_t gc_launch_encode(int64_t n, int64_t ld, int n_cells, uint32_t radix, const int8_t *cells,
                             uint32_t *index, cudaStream_t st)
{
    const int64_t need = (n + kThreads * kEPT - 1) / (kThreads * kEPT);
    encode_kernel<<<static_cast<int>(need < 1 ? 1 : (need > 65535 ? 65535 : need)), kThreads, 0, st>>>(
        n, ld, n_cells, radix, cells, index);
    return cudaGetLastError();
}

cudaError_t gc_launch_decode(int64_t n, int64_t ld, int n_cells, uint32_t radix, const uint32_t *index,
                             int8_t *cells, cudaStream_t st)
{
    const int64_t need = (n + kThreads * kEPT - 1) / (kThreads * kEPT);
    decode_kernel<<<static_cast<int>(need < 1 ? 1 : (need > 65535 ? 65535 : need)), kThreads, 0, st>>>(
        n, ld, n_cells, radix, index, cells);
    return cudaGetLastError();
}
