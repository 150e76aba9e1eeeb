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
// Cellular (polarisation family) step: C cells, per-cell identical [S][A] tables.
template <int C, int RNG>
__global__ void __launch_bounds__(kThreads)
cell_step_kernel(const __grid_constant__ CellTables tab, const __grid_constant__ StepIO io)
{
    __shared__ uint2 s_sa[GC_TBL];                 // .x packed move/noisy/draws, .y reward bits
    __shared__ float s_rn[GC_TBL];                 // reward when the draw fired
    __shared__ uint8_t s_se[C][GC_TBL];
    __shared__ unsigned long long s_stats[5];

    for (int i = threadIdx.x; i < GC_TBL; i += kThreads)
        s_sa[i] = make_uint2(tab.sa[i], __float_as_uint(tab.reward[i])), s_rn[i] = tab.reward_noisy[i];
    for (int i = threadIdx.x; i < C * GC_TBL; i += kThreads)
        s_se[i / GC_TBL][i % GC_TBL] = tab.se[i / GC_TBL][i % GC_TBL];
    if (threadIdx.x < 5) s_stats[threadIdx.x] = 0;
    __syncthreads();

    const uint32_t step_counter = launch_step_counter(io);
    ThreadStats ts = {0, 0, 0, 0, 0};
    const int64_t ld = io.ld;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads * kEPT;
    for (int64_t e0 = io.begin + (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * kEPT;
         e0 < io.end; e0 += stride) {
        uint32_t sw[C], aw[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            sw[c] = ld_stream_u32(io.state + c * ld + e0);
            aw[c] = ld_stream_u32(io.actions + c * ld + e0);
        }
        const int4 t4 = ld_stream_v4(io.t + e0);
        const int tin[kEPT] = {t4.x, t4.y, t4.z, t4.w};

        uint32_t nsw[C], sew[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { nsw[c] = 0; sew[c] = 0; }
        int tout[kEPT];
        float rout[kEPT];
        uint32_t iout[kEPT];
        uint32_t trunc_w = 0, unsafe_w = 0, count_w = 0;

#pragma unroll
        for (int e = 0; e < kEPT; ++e) {
            const bool valid = (e0 + e) < io.end;
            uint32_t ns[C];
            float r = 0.0f;
            uint32_t rnd[4] = {0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const uint32_t s = byte_of(sw[c], e), a = byte_of(aw[c], e);
                const uint32_t sa_ix = (s * GC_LVL_PAD + a) & (GC_TBL - 1);
                const uint2 ent = s_sa[sa_ix];
                uint32_t nxt = ent.x & 15u;
                bool fire = false;
                if (RNG == GC_RNG_PHILOX) {
                    if ((c & 3) == 0) {
                        const uint64_t gid = static_cast<uint64_t>(io.env_id_offset + e0 + e);
                        const uint32_t ctr = io.episodic ? static_cast<uint32_t>(tin[e]) : step_counter;
                        philox4x32_10(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), ctr,
                                      static_cast<uint32_t>(c >> 2), io.round_key, rnd);
                    }
                    fire = (ent.x & 0x100u) && (static_cast<unsigned long long>(rnd[c & 3]) < tab.noise_thr);
                } else if (RNG == GC_RNG_REPLAY) {
                    if (valid && (ent.x & 0x100u)) fire = io.replay[(e0 + e) * C + c] < tab.noise_prob;
                }
                if (fire) nxt = (ent.x >> 4) & 15u;
                r += fire ? s_rn[sa_ix] : __uint_as_float(ent.y);             // left to right, from 0.0
                ns[c] = nxt;
            }
            if (tab.reward_log2) r = log1pf(r) * 1.44269504088896341f;

            // row 0 of the side-effects matrix: entry j from (s'_0, s'_p), p = 1 for j = 0
            uint32_t uns = 0, cnt = 0, idx = 0;
            uint32_t code[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const uint32_t partner = (c == 0) ? ns[C > 1 ? 1 : 0] : ns[c];
                code[c] = s_se[c][(ns[0] * GC_LVL_PAD + partner) & (GC_TBL - 1)];
                uns |= (code[c] == 2u);
                cnt += (tab.counted_mask >> ns[c]) & 1u;
                idx += ns[c] * tab.place[c];
            }
            int tn = tin[e] + 1;
            uint32_t tr = 0;
            if (io.max_episode_steps > 0 && tn >= io.max_episode_steps) {   // fused time-limit auto-reset
                tr = 1; tn = 0; idx = tab.init_index;
#pragma unroll
                for (int c = 0; c < C; ++c) ns[c] = static_cast<uint32_t>(tab.init[c]);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                nsw[c] |= ns[c] << (8 * e);
                sew[c] |= code[c] << (8 * e);
            }
            tout[e] = tn; rout[e] = r; iout[e] = idx;
            trunc_w |= tr << (8 * e); unsafe_w |= uns << (8 * e); count_w |= cnt << (8 * e);
            if (valid) {
                ts.steps += 1; ts.unsafe += uns; ts.count += cnt; ts.truncated += tr;
                ts.reward_q24 += __float2ll_rn(r * 16777216.0f);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) st_stream_u32(io.state + c * ld + e0, nsw[c]);
        if (io.se_row) {
#pragma unroll
            for (int c = 0; c < C; ++c) st_stream_u32(io.se_row + c * ld + e0, sew[c]);
        }
        st_stream_v4(io.t + e0, make_int4(tout[0], tout[1], tout[2], tout[3]));
        st_stream_v4(io.reward + e0, make_int4(__float_as_int(rout[0]), __float_as_int(rout[1]),
                                               __float_as_int(rout[2]), __float_as_int(rout[3])));
        st_stream_v4(io.index + e0, make_int4(iout[0], iout[1], iout[2], iout[3]));
        st_stream_u32(io.terminated + e0, 0u);
        st_stream_u32(io.truncated + e0, trunc_w);
        st_stream_u32(io.unsafe + e0, unsafe_w);
        st_stream_u32(io.count + e0, count_w);
    }
    if (io.stats) block_flush_stats(ts, s_stats, io.stats);
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

template <int C>
cudaError_t launch_cell_c(const CellTables &tab, const StepIO &io, int rng_mode, int n_sm, cudaStream_t st)
{
    const int64_t n = io.end - io.begin;
    switch (rng_mode) {
    case GC_RNG_NONE:
        cell_step_kernel<C, GC_RNG_NONE><<<grid_for<cell_step_kernel<C, GC_RNG_NONE>>(n, n_sm), kThreads, 0, st>>>(tab, io);
        break;
    case GC_RNG_PHILOX:
        cell_step_kernel<C, GC_RNG_PHILOX><<<grid_for<cell_step_kernel<C, GC_RNG_PHILOX>>(n, n_sm), kThreads, 0, st>>>(tab, io);
        break;
    default:
        cell_step_kernel<C, GC_RNG_REPLAY><<<grid_for<cell_step_kernel<C, GC_RNG_REPLAY>>(n, n_sm), kThreads, 0, st>>>(tab, io);
        break;
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t gc_launch_cell_step(const CellTables &tab, const StepIO &io, int rng_mode, int n_sm, cudaStream_t st)
{
    switch (tab.n_cells) {
#define GC_CASE(C) case C: return launch_cell_c<C>(tab, io, rng_mode, n_sm, st);
        GC_CASE(1) GC_CASE(2) GC_CASE(3) GC_CASE(4) GC_CASE(5) GC_CASE(6) GC_CASE(7) GC_CASE(8)
        GC_CASE(9) GC_CASE(10) GC_CASE(11) GC_CASE(12) GC_CASE(13) GC_CASE(14) GC_CASE(15) GC_CASE(16)
#undef GC_CASE
    default: return cudaErrorInvalidValue;
    }
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
