// Chunked parallel-in-time Viterbi with on-device parallel traceback
// (hmm.py:668-676 -> _hmm.pyx:201-259).
//
// DP: one warp per chunk, lane j owns column j of log A in registers, delta is
// kept max-normalised (max = 0) and broadcast through shared memory; the
// strict-'>' / lowest-from-state tie rule of the reference is kept.  Chunks
// that do not begin a sequence speculate their start vector with `warmup`
// steps from a flat vector and are verified / repaired like forward.cu.
//
// Traceback without a serial walk over T: while the DP runs, each lane also
// carries, per traceback tile of 64 steps, the state at the end of the
// PREVIOUS tile that its best path comes from (one SHFL per step).  Those
// tile maps are composed per chunk, the per-chunk maps are walked once per
// sequence out of shared memory, and then every tile is traced independently
// (one lane per tile) through the uint8 back-pointers.
//
// The returned log-probability is a float64 re-score of the returned path
// against the float64 tables, so it does not carry fp32 DP rounding.
//
// Algorithmic HBM bytes per step: 4N (elog) read, NP (back-pointers) written,
// then ~NP read + 1 written by the traceback.
#include "scan.cuh"

#define VIT_U 4

template <typename T, int NS, bool RATIO>
__global__ void __launch_bounds__(TEHMM_WARPS_PER_CTA * 32, (sizeof(T) == 4 && NS == 1) ? 3 : 1)
viterbi_kernel(TehmmModelDev m, TehmmBatchDev b, const T *__restrict__ elog,
               const double *__restrict__ ratios, uint8_t *__restrict__ bp,
               uint8_t *__restrict__ tilemap, T *__restrict__ start_vec,
               T *__restrict__ end_vec, const int *__restrict__ bad, int mode)
{
    constexpr int NP = 32 * NS;
    __shared__ __align__(16) T ds_all[TEHMM_WARPS_PER_CTA][2][NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T(*ds)[NP] = ds_all[warp];
    const int N = m.N;
    const T NEG = (T)-INFINITY;

    T c[NS][NP];
    T ls[NS], dg[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int j = lane + 32 * s;
#pragma unroll
        for (int i = 0; i < NP; ++i) c[s][i] = (T)m.cut_trans[(int64_t)i * NP + j];
        ls[s] = (T)m.cut_start[j];
        dg[s] = (T)m.cut_trans[(int64_t)j * NP + j];
    }
    const T a00 = (T)m.cut_trans[0];

    for (int64_t ci = (int64_t)blockIdx.x * TEHMM_WARPS_PER_CTA + warp; ci < b.nchunks;
         ci += (int64_t)gridDim.x * TEHMM_WARPS_PER_CTA) {
        if (mode == 1 && !bad[ci]) continue;
        const TehmmChunk ch = b.chunks[ci];
        T d[NS];
        int orig[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) orig[s] = 0;
        int buf = 0;

        auto load_e = [&](int64_t t, T (&et)[NS]) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                int j = lane + 32 * s;
                et[s] = j < N ? elog[t * N + j] : NEG;
            }
        };
        auto init = [&](int64_t t, const T (&et)[NS]) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                d[s] = ls[s] + et[s];
                if (RATIO) {
                    double r = ratios[t];
                    if (r > 1.0) d[s] += dg[s] * (T)(r - 1.0);
                }
            }
            log_normalise<T, NS>(d);
        };
        // one DP step; arg[] receives the back-pointers
        auto step = [&](int64_t t, const T (&et)[NS], int (&arg)[NS]) {
#pragma unroll
            for (int s = 0; s < NS; ++s) ds[buf][lane + 32 * s] = d[s];
            __syncwarp();
            T extra0[NS], add[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) { extra0[s] = (T)0; add[s] = et[s]; }
            if (RATIO) {
                // _hmm.pyx:234-237 vs 243-244: from-state 0 always gets A_jj*r
                // (minus A_00 when j==0); the others get A_jj*(r-1) only if r>1
                const T r = (T)ratios[t];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const T common = r > (T)1 ? dg[s] * (r - (T)1) : (T)0;
                    add[s] += common;
                    // (a zero-probability self transition would give inf-inf here; the
                    //  reference's finite -1e100 sentinel has no fp32 twin)
                    extra0[s] = dg[s] > NEG ? dg[s] * r - common : (T)0;
                    if (lane + 32 * s == 0 && a00 > NEG) extra0[s] -= a00;
                }
            }
            T y[NS];
            matvec_max<T, NS>(ds[buf], c, extra0, y, arg);
            buf ^= 1;
#pragma unroll
            for (int s = 0; s < NS; ++s) d[s] = y[s] + add[s];
            log_normalise<T, NS>(d);
        };

        int64_t t = ch.t0;
        if (ch.t0 > ch.s0) {
            if (mode == 0) {
                int64_t tw = ch.t0 - b.warmup;
                T et[NS];
                int arg[NS];
                if (tw <= ch.s0) {
                    tw = ch.s0;
                    load_e(tw, et);
                    init(tw, et);
                    ++tw;
                } else {
#pragma unroll
                    for (int s = 0; s < NS; ++s) d[s] = (lane + 32 * s) < N ? (T)0 : NEG;
                }
                for (; tw < ch.t0; ++tw) {
                    load_e(tw, et);
                    step(tw, et, arg);
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) start_vec[ci * NP + lane + 32 * s] = d[s];
            } else {
#pragma unroll
                for (int s = 0; s < NS; ++s) d[s] = start_vec[ci * NP + lane + 32 * s];
            }
        } else {
            T et[NS];
            load_e(t, et);
            init(t, et);
            ++t;
        }

        T en[VIT_U][NS];
#pragma unroll
        for (int u = 0; u < VIT_U; ++u) load_e(min(t + u, ch.t1 - 1), en[u]);
        for (; t < ch.t1; t += VIT_U) {
            T ec[VIT_U][NS];
#pragma unroll
            for (int u = 0; u < VIT_U; ++u) {
#pragma unroll
                for (int s = 0; s < NS; ++s) ec[u][s] = en[u][s];
            }
#pragma unroll
            for (int u = 0; u < VIT_U; ++u) load_e(min(t + VIT_U + u, ch.t1 - 1), en[u]);
#pragma unroll
            for (int u = 0; u < VIT_U; ++u) {
                const int64_t tc = t + u;
                if (tc >= ch.t1) continue;
                int arg[NS];
                step(tc, ec[u], arg);
                const int rel = (int)((tc - ch.s0) & (TEHMM_TILE - 1));
#pragma unroll
                for (int s = 0; s < NS; ++s) bp[tc * NP + lane + 32 * s] = (uint8_t)arg[s];
                // carry "state at the end of the previous tile" along the best paths
                int no[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (rel == 0) {
                        no[s] = arg[s];
                    } else if (NS == 1) {
                        no[s] = __shfl_sync(TEHMM_FULL, orig[0], arg[s]);
                    } else {
                        int lo = __shfl_sync(TEHMM_FULL, orig[0], arg[s] & 31);
                        int hi = __shfl_sync(TEHMM_FULL, orig[NS - 1], arg[s] & 31);
                        no[s] = arg[s] < 32 ? lo : hi;
                    }
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) orig[s] = no[s];
                if (rel == TEHMM_TILE - 1 || tc == ch.t1 - 1) {
                    const int64_t tile = ch.tile0 + ((tc - ch.t0) >> 6);
#pragma unroll
                    for (int s = 0; s < NS; ++s) tilemap[tile * NP + lane + 32 * s] = (uint8_t)orig[s];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) end_vec[ci * NP + lane + 32 * s] = d[s];
    }
}

// cmap[c][j] = state at t0-1 on the best path that ends chunk c in state j
__global__ void vit_compose_kernel(TehmmBatchDev b, int NP, const uint8_t *__restrict__ tilemap,
                                   uint8_t *__restrict__ cmap)
{
    int64_t ci = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (ci >= b.nchunks) return;
    const TehmmChunk ch = b.chunks[ci];
    for (int j = lane; j < NP; j += 32) {
        int x = j;
        for (int k = ch.ntiles - 1; k >= 0; --k) x = tilemap[(ch.tile0 + k) * NP + x];
        cmap[ci * NP + j] = (uint8_t)x;
    }
}

// One CTA per sequence: end state of the last chunk = first maximum of its
// final delta (np.argmax, _hmm.pyx:252); then walk the chunk maps backwards out
// of shared memory to get every chunk's end state.
template <typename T>
__global__ void vit_seqscan_kernel(TehmmBatchDev b, int NP, const T *__restrict__ end_vec,
                                   const uint8_t *__restrict__ cmap, uint8_t *__restrict__ chunk_end,
                                   int smem_chunks)
{
    extern __shared__ uint8_t cm[];
    __shared__ int cur;
    int64_t s = blockIdx.x;
    int64_t c0 = b.seq_chunk0[s], c1 = b.seq_chunk0[s + 1];
    if (c1 <= c0) return;
    if (threadIdx.x == 0) {
        const T *dv = end_vec + (c1 - 1) * NP;
        int best = 0;
        for (int j = 1; j < NP; ++j)
            if (dv[j] > dv[best]) best = j;
        cur = best;
        chunk_end[c1 - 1] = (uint8_t)best;
    }
    __syncthreads();
    for (int64_t hi = c1; hi > c0 + 1; hi -= smem_chunks) {
        int64_t lo = hi - smem_chunks;
        if (lo < c0 + 1) lo = c0 + 1;
        // stage cmap[lo..hi)
        int64_t bytes = (hi - lo) * NP;
        for (int64_t e = threadIdx.x * 16; e < bytes; e += (int64_t)blockDim.x * 16)
            *reinterpret_cast<uint4 *>(cm + e) = *reinterpret_cast<const uint4 *>(cmap + lo * NP + e);
        __syncthreads();
        if (threadIdx.x == 0) {
            int st = cur;
            for (int64_t c = hi - 1; c >= lo; --c) {
                st = cm[(c - lo) * NP + st];
                chunk_end[c - 1] = (uint8_t)st;
            }
            cur = st;
        }
        __syncthreads();
    }
}

// One warp per chunk: lane 0 resolves the end state of every tile, then the
// lanes trace the tiles in parallel.
__global__ void __launch_bounds__(TEHMM_WARPS_PER_CTA * 32)
vit_traceback_kernel(TehmmBatchDev b, int NP, const uint8_t *__restrict__ bp,
                     const uint8_t *__restrict__ tilemap, const uint8_t *__restrict__ chunk_end,
                     uint8_t *__restrict__ states, int64_t *__restrict__ states64, int max_tiles)
{
    extern __shared__ uint8_t tile_end_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *tile_end = tile_end_all + (size_t)warp * max_tiles;
    for (int64_t ci = (int64_t)blockIdx.x * TEHMM_WARPS_PER_CTA + warp; ci < b.nchunks;
         ci += (int64_t)gridDim.x * TEHMM_WARPS_PER_CTA) {
        const TehmmChunk ch = b.chunks[ci];
        __syncwarp();
        if (lane == 0) {
            int st = chunk_end[ci];
            for (int k = ch.ntiles - 1; k >= 0; --k) {
                tile_end[k] = (uint8_t)st;
                st = tilemap[(ch.tile0 + k) * NP + st];
            }
        }
        __syncwarp();
        for (int k = lane; k < ch.ntiles; k += 32) {
            const int64_t ta = ch.t0 + (int64_t)k * TEHMM_TILE;
            int64_t tz = ta + TEHMM_TILE;
            if (tz > ch.t1) tz = ch.t1;
            int st = tile_end[k];
            for (int64_t t = tz - 1; t >= ta; --t) {
                if (states) states[t] = (uint8_t)st;
                if (states64) states64[t] = st;
                if (t > ta) st = bp[t * NP + st];
            }
        }
    }
}

// float64 score of the returned path, in the reference's terms
// (_hmm.pyx:222-225, 232-248): one warp per chunk, lanes over time.
template <typename OBS>
__global__ void vit_rescore_kernel(TehmmModelDev m, TehmmBatchDev b, const OBS *__restrict__ obs,
                                   const uint8_t *__restrict__ states,
                                   const double *__restrict__ ratios_em,
                                   const double *__restrict__ ratios_dp,
                                   double *__restrict__ score_part)
{
    int64_t ci = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (ci >= b.nchunks) return;
    const TehmmChunk ch = b.chunks[ci];
    const int N = m.N;
    double acc = 0.0;
    for (int64_t t = ch.t0 + lane; t < ch.t1; t += 32) {
        const int j = states[t];
        double e = 0.0;
        for (int k = 0; k < m.K; ++k)
            e += m.table[((int64_t)k * N + j) * m.S + (int64_t)obs[t * m.K + k]];
        e *= m.normalize;
        if (ratios_em) e *= ratios_em[t];
        const double ajj = m.log_trans[(int64_t)j * N + j];
        double term;
        if (t == ch.s0) {
            term = m.log_start[j] + e;
            if (ratios_dp && ratios_dp[t] > 1.) term += ajj * (ratios_dp[t] - 1.);
        } else {
            const int i = states[t - 1];
            term = m.log_trans[(int64_t)i * N + j] + e;
            if (ratios_dp) {
                const double r = ratios_dp[t];
                if (i == 0) {
                    term += ajj * r;
                    if (j == 0) term -= m.log_trans[0];
                } else if (r > 1.) {
                    term += ajj * (r - 1.);
                }
            }
        }
        acc += term;
    }
    acc = warp_sum(acc);
    if (lane == 0) score_part[ci] = acc;
}

__global__ void vit_score_reduce_kernel(TehmmBatchDev b, const double *__restrict__ score_part,
                                        double *__restrict__ logprob)
{
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= b.nseq) return;
    double acc = 0.0;
    for (int64_t c = b.seq_chunk0[s] + lane; c < b.seq_chunk0[s + 1]; c += 32) acc += score_part[c];
    acc = warp_sum(acc);
    if (lane == 0) logprob[s] = acc;
}

template <typename T, int NS>
static cudaError_t launch_vit(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                              const T *elog, const double *ratios, uint8_t *bp, uint8_t *tilemap,
                              T *start_vec, T *end_vec, const int *bad, int mode, int grid)
{
    if (ratios)
        viterbi_kernel<T, NS, true><<<grid, TEHMM_WARPS_PER_CTA * 32, 0, st>>>(m, b, elog, ratios, bp, tilemap, start_vec, end_vec, bad, mode);
    else
        viterbi_kernel<T, NS, false><<<grid, TEHMM_WARPS_PER_CTA * 32, 0, st>>>(m, b, elog, ratios, bp, tilemap, start_vec, end_vec, bad, mode);
    return cudaGetLastError();
}

cudaError_t tehmm_launch_viterbi(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                                 int prec, const void *elog, const double *ratios, uint8_t *bp,
                                 uint8_t *tilemap, void *start_vec, void *end_vec, const int *bad,
                                 int mode, int grid)
{
    if (prec == TEHMM_F32) {
        if (m.NS == 1) return launch_vit<float, 1>(st, m, b, (const float *)elog, ratios, bp, tilemap, (float *)start_vec, (float *)end_vec, bad, mode, grid);
        return launch_vit<float, 2>(st, m, b, (const float *)elog, ratios, bp, tilemap, (float *)start_vec, (float *)end_vec, bad, mode, grid);
    }
    if (m.NS == 1) return launch_vit<double, 1>(st, m, b, (const double *)elog, ratios, bp, tilemap, (double *)start_vec, (double *)end_vec, bad, mode, grid);
    return launch_vit<double, 2>(st, m, b, (const double *)elog, ratios, bp, tilemap, (double *)start_vec, (double *)end_vec, bad, mode, grid);
}

// compose + per-sequence scan + per-tile traceback + fp64 re-score: 5 launches
cudaError_t tehmm_launch_traceback(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                                   int prec, const void *end_vec, const uint8_t *bp,
                                   const uint8_t *tilemap, uint8_t *cmap, uint8_t *chunk_end,
                                   uint8_t *states, int64_t *states64,
                                   const double *ratios_em, const double *ratios_dp,
                                   double *score_part, double *logprob, int max_tiles, int grid)
{
    const int NP = m.NP;
    int warps = 4;
    vit_compose_kernel<<<(int)((b.nchunks + warps - 1) / warps), warps * 32, 0, st>>>(b, NP, tilemap, cmap);
    int smem_chunks = (160 * 1024) / NP;
    size_t smem = (size_t)smem_chunks * NP;
    if (prec == TEHMM_F32) {
        cudaFuncSetAttribute(vit_seqscan_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        vit_seqscan_kernel<float><<<(int)b.nseq, 256, smem, st>>>(b, NP, (const float *)end_vec, cmap, chunk_end, smem_chunks);
    } else {
        cudaFuncSetAttribute(vit_seqscan_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        vit_seqscan_kernel<double><<<(int)b.nseq, 256, smem, st>>>(b, NP, (const double *)end_vec, cmap, chunk_end, smem_chunks);
    }
    size_t tsm = (size_t)TEHMM_WARPS_PER_CTA * max_tiles;
    vit_traceback_kernel<<<grid, TEHMM_WARPS_PER_CTA * 32, tsm, st>>>(b, NP, bp, tilemap, chunk_end, states, states64, max_tiles);
    int rgrid = (int)((b.nchunks + warps - 1) / warps);
    if (b.obs_bytes == 1) vit_rescore_kernel<uint8_t><<<rgrid, warps * 32, 0, st>>>(m, b, (const uint8_t *)b.obs, states, ratios_em, ratios_dp, score_part);
    else if (b.obs_bytes == 2) vit_rescore_kernel<uint16_t><<<rgrid, warps * 32, 0, st>>>(m, b, (const uint16_t *)b.obs, states, ratios_em, ratios_dp, score_part);
    else vit_rescore_kernel<int32_t><<<rgrid, warps * 32, 0, st>>>(m, b, (const int32_t *)b.obs, states, ratios_em, ratios_dp, score_part);
    vit_score_reduce_kernel<<<(int)((b.nseq + warps - 1) / warps), warps * 32, 0, st>>>(b, score_part, logprob);
    return cudaGetLastError();
}
