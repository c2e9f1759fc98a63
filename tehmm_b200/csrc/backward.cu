// Chunked parallel-in-time backward pass fused with the posterior glue and the
// Baum-Welch transition counts (hmm.py:715-729,545-568; basehmm.py:265-272,
// 516-517; _hmm.pyx:62-117,160-198).
//
// Same speculate / verify / repair scheme as forward.cu, run right-to-left.
// Lane i owns ROW i of the transition matrix.  At step t the warp holds
//   w[j]    = b_{t+1}[j] * beta_hat_{t+1}[j]            (broadcast through smem)
//   beta'[i]= sum_j A[i][j] w[j]
//   Z       = sum_i alpha_hat_t[i] beta'[i]
//   gamma_t = alpha_hat_t .* beta' / Z                  (posterior, rows sum to 1)
//   xi_t    = alpha_hat_t[i] A[i][j] w[j] / Z           (sums to 1 over i,j)
// so neither the T x N x N lneta tensor nor a T x N beta lattice is ever
// written.  The factor A[i][j] and the reference's 1/N (its beta[T-1] is
// log(1/N), _hmm.pyx:179) are applied once, in the reduction kernel.
//
// Algorithmic HBM bytes per step (fp32): 4N (b) + 4N (alpha) read,
// + 4N written when posteriors are requested, + 1 for MAP states.
#include "scan.cuh"

#define BWD_U 4

template <typename T, int NS, bool RATIO, bool TRANS>
__global__ void __launch_bounds__(TEHMM_WARPS_PER_CTA * 32, (sizeof(T) == 4 && NS == 1 && !TRANS) ? 3 : 1)
backward_kernel(TehmmModelDev m, TehmmBatchDev b, int flags, const T *__restrict__ blin,
                const T *__restrict__ alpha, const double *__restrict__ ratios,
                T *__restrict__ post, uint8_t *__restrict__ map_states,
                double *__restrict__ map_part, T *__restrict__ xi_part,
                T *__restrict__ xdiag_part, T *__restrict__ gamma0, T *__restrict__ start_vec,
                T *__restrict__ end_vec, const int *__restrict__ bad, int mode)
{
    constexpr int NP = 32 * NS;
    __shared__ __align__(16) T ws_all[TEHMM_WARPS_PER_CTA][2][NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T(*ws)[NP] = ws_all[warp];
    const int N = m.N;
    const bool want_post = (flags & TEHMM_BWD_POSTERIORS) != 0;
    const bool want_map = (flags & TEHMM_BWD_MAP) != 0;
    const bool renorm = (flags & TEHMM_BWD_RENORM_EPS) != 0;
    const double eps32 = 1.1920928955078125e-07;
    const double renorm_den = 1.0 + (double)N * eps32;

    // row i of the transition matrix for each owned state
    T c[NS][NP];
    double dg[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        int i = lane + 32 * s;
#pragma unroll
        for (int j = 0; j < NP; ++j) c[s][j] = (T)m.lin_trans[(int64_t)i * NP + j];
        dg[s] = RATIO ? m.cut_trans[(int64_t)i * NP + i] : 0.0;
    }

    for (int64_t ci = (int64_t)blockIdx.x * TEHMM_WARPS_PER_CTA + warp; ci < b.nchunks;
         ci += (int64_t)gridDim.x * TEHMM_WARPS_PER_CTA) {
        if (mode == 1 && !bad[ci]) continue;
        const TehmmChunk ch = b.chunks[ci];
        T u[NS];                 // canonical beta_hat at the step last processed
        int buf = 0;
        T xi[TRANS ? NS : 1][TRANS ? NP : 1];
        T xd[NS];
        if (TRANS) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                xd[s] = (T)0;
#pragma unroll
                for (int j = 0; j < NP; ++j) xi[s][j] = (T)0;
            }
        }
        double mapsum = 0.0;

        // w_{t+1} = b_{t+1} .* u [.* g_{t+1}] -> smem ; returns beta'_t (unnormalised)
        auto beta_step = [&](int64_t t, const T (&bt1)[NS], T (&bp)[NS], T (&wv)[NP], bool keep) {
            T w[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) w[s] = bt1[s] * u[s];
            if (RATIO) {
                double r = ratios[t + 1];
                if (r > 1.0) {
                    double lg[NS], mg = -INFINITY;
#pragma unroll
                    for (int s = 0; s < NS; ++s) { lg[s] = dg[s] * (r - 1.0); mg = fmax(mg, lg[s]); }
                    mg = warp_max(mg);
#pragma unroll
                    for (int s = 0; s < NS; ++s) w[s] = mg > -INFINITY ? w[s] * (T)exp(lg[s] - mg) : (T)0;
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) ws[buf][lane + 32 * s] = w[s];
            __syncwarp();
            if (keep) matvec_sum_keep<T, NS>(ws[buf], c, bp, wv);
            else matvec_sum<T, NS>(ws[buf], c, bp);
            buf ^= 1;
        };
        auto load_row = [&](const T *src, int64_t t, T (&v)[NS]) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                int j = lane + 32 * s;
                v[s] = j < N ? src[t * N + j] : (T)0;
            }
        };

        // ---- phase A: beta_hat at t1 (speculated), unless the chunk ends its sequence
        if (ch.t1 < ch.s1) {
            if (mode == 0) {
                int64_t tq = ch.t1 + b.warmup;
                if (tq > ch.s1 - 1) tq = ch.s1 - 1;
#pragma unroll
                for (int s = 0; s < NS; ++s) u[s] = (lane + 32 * s) < N ? (T)1 : (T)0;
                for (int64_t t = tq - 1; t >= ch.t1; --t) {
                    T bt1[NS], bp[NS], dummy[NP];
                    load_row(blin, t + 1, bt1);
                    beta_step(t, bt1, bp, dummy, false);
                    canonicalise<T, NS>(bp);
#pragma unroll
                    for (int s = 0; s < NS; ++s) u[s] = bp[s];
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) start_vec[ci * NP + lane + 32 * s] = u[s];
            } else {
#pragma unroll
                for (int s = 0; s < NS; ++s) u[s] = start_vec[ci * NP + lane + 32 * s];
            }
        }

        // ---- phase B: t = t1-1 ... t0 with outputs
        T an[BWD_U][NS], bn[BWD_U][NS];
        int64_t t = ch.t1 - 1;
#pragma unroll
        for (int q = 0; q < BWD_U; ++q) {
            int64_t tt = max(t - q, ch.t0);
            load_row(alpha, tt, an[q]);
            load_row(blin, min(tt + 1, ch.s1 - 1), bn[q]);
        }
        for (; t >= ch.t0; t -= BWD_U) {
            T ac[BWD_U][NS], bc[BWD_U][NS];
#pragma unroll
            for (int q = 0; q < BWD_U; ++q) {
#pragma unroll
                for (int s = 0; s < NS; ++s) { ac[q][s] = an[q][s]; bc[q][s] = bn[q][s]; }
            }
#pragma unroll
            for (int q = 0; q < BWD_U; ++q) {
                int64_t tt = max(t - BWD_U - q, ch.t0);
                load_row(alpha, tt, an[q]);
                load_row(blin, min(tt + 1, ch.s1 - 1), bn[q]);
            }
#pragma unroll
            for (int q = 0; q < BWD_U; ++q) {
                const int64_t tc = t - q;
                if (tc < ch.t0) continue;
                T bp[NS];
                T wv[TRANS ? NP : 1];
                const bool last = (tc == ch.s1 - 1);
                if (last) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) bp[s] = (lane + 32 * s) < N ? (T)1 : (T)0;
                } else if (TRANS) {
                    T(&wref)[NP] = reinterpret_cast<T(&)[NP]>(wv);
                    beta_step(tc, bc[q], bp, wref, true);
                } else {
                    T dummy[NP];
                    beta_step(tc, bc[q], bp, dummy, false);
                }
                // posterior
                T p[NS], zl = (T)0;
#pragma unroll
                for (int s = 0; s < NS; ++s) { p[s] = ac[q][s] * bp[s]; zl += p[s]; }
                const T Z = warp_sum(zl);
                const T invZ = (T)1 / Z;
                T g[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) g[s] = p[s] * invZ;
                if (TRANS) {
                    if (!last) {
                        T(&wref)[NP] = reinterpret_cast<T(&)[NP]>(wv);
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            const T qv = ac[q][s] * invZ;
#pragma unroll
                            for (int j = 0; j < NP; ++j) xi[s][j] = fma(qv, wref[j], xi[s][j]);
                        }
                    }
                    if (RATIO && tc > ch.s0) {
                        // implied self transitions of a long segment (_hmm.pyx:89-96,106-111)
                        double r = ratios[tc];
                        if (r > 1.0) {
#pragma unroll
                            for (int s = 0; s < NS; ++s) xd[s] += (T)(r - 1.0) * g[s];
                        }
                    }
                    if (tc == ch.s0) {
#pragma unroll
                        for (int s = 0; s < NS; ++s) gamma0[(int64_t)ch.seq * NP + lane + 32 * s] = g[s];
                    }
                }
                if (want_post) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        int j = lane + 32 * s;
                        if (j < N) {
                            T gv = g[s];
                            if (renorm) gv = (T)(((double)gv + eps32) / renorm_den);
                            post[tc * N + j] = gv;
                        }
                    }
                }
                if (want_map) {
                    // argmax with the lowest state winning ties (np.argmax, basehmm.py:357)
                    T best;
                    int arg;
                    if (sizeof(T) == 4 && NS == 1) {
                        // posteriors are >= 0: their bit patterns order like the values
                        const unsigned bits = lane < N ? __float_as_uint((float)g[0]) : 0u;
                        const unsigned mx = __reduce_max_sync(TEHMM_FULL, bits);
                        arg = __ffs(__ballot_sync(TEHMM_FULL, bits == mx)) - 1;
                        best = (T)__uint_as_float(mx);
                    } else {
                        best = g[0];
                        arg = lane;
#pragma unroll
                        for (int s = 1; s < NS; ++s)
                            if (g[s] > best) { best = g[s]; arg = lane + 32 * s; }
                        if (arg >= N) best = (T)-1;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            T ob = __shfl_xor_sync(TEHMM_FULL, best, o);
                            int oa = __shfl_xor_sync(TEHMM_FULL, arg, o);
                            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
                        }
                    }
                    if (lane == 0) {
                        map_states[tc] = (uint8_t)arg;
                        mapsum += renorm ? ((double)best + eps32) / renorm_den : (double)best;
                    }
                }
                canonicalise<T, NS>(bp);
#pragma unroll
                for (int s = 0; s < NS; ++s) u[s] = bp[s];
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) end_vec[ci * NP + lane + 32 * s] = u[s];
        if (want_map && lane == 0) map_part[ci] = mapsum;
        if (TRANS) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                T *dst = xi_part + ((int64_t)ci * NP + lane + 32 * s) * NP;
#pragma unroll
                for (int j = 0; j < NP; j += 4) {
                    if (sizeof(T) == 4)
                        *reinterpret_cast<float4 *>(dst + j) = make_float4(xi[s][j], xi[s][j + 1], xi[s][j + 2], xi[s][j + 3]);
                    else { dst[j] = xi[s][j]; dst[j + 1] = xi[s][j + 1]; dst[j + 2] = xi[s][j + 2]; dst[j + 3] = xi[s][j + 3]; }
                }
                xdiag_part[(int64_t)ci * NP + lane + 32 * s] = xd[s];
            }
        }
    }
}

// start[i] += sum_seq gamma0[seq][i];
// trans[i][j] += (1/N) * ( A[i][j] * sum_chunks xi_part[c][i][j] + [i==j] sum_chunks xdiag_part[c][i] )
// Deterministic: fixed summation order, float64 accumulation.
template <typename T>
__global__ void trans_reduce_kernel(TehmmModelDev m, TehmmBatchDev b, const T *__restrict__ xi_part,
                                    const T *__restrict__ xdiag_part, const T *__restrict__ gamma0,
                                    double *__restrict__ start_trans)
{
    const int N = m.N, NP = m.NP;
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < N) {
        double acc = 0.0;
        for (int64_t s = 0; s < b.nseq; ++s)
            if (b.seq_off[s + 1] > b.seq_off[s]) acc += (double)gamma0[s * NP + e];
        start_trans[e] += acc;
    }
    if (e < N * N) {
        int i = e / N, j = e - i * N;
        double acc = 0.0, dacc = 0.0;
        for (int64_t c = 0; c < b.nchunks; ++c) {
            acc += (double)xi_part[(c * NP + i) * NP + j];
            if (i == j) dacc += (double)xdiag_part[c * NP + i];
        }
        start_trans[N + e] += (acc * m.lin_trans[(int64_t)i * NP + j] + dacc) / (double)N;
    }
}

// map_score[seq] = sum over the sequence's chunks (basehmm.py:358)
__global__ void map_reduce_kernel(TehmmBatchDev b, const double *__restrict__ map_part,
                                  double *__restrict__ map_score)
{
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= b.nseq) return;
    double acc = 0.0;
    for (int64_t c = b.seq_chunk0[s] + lane; c < b.seq_chunk0[s + 1]; c += 32) acc += map_part[c];
    acc = warp_sum(acc);
    if (lane == 0) map_score[s] = acc;
}

template <typename T, int NS>
static cudaError_t launch_bwd(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                              int flags, const T *blin, const T *alpha, const double *ratios,
                              T *post, uint8_t *map_states, double *map_part, T *xi_part,
                              T *xdiag_part, T *gamma0, T *start_vec, T *end_vec, const int *bad,
                              int mode, int grid)
{
    const int th = TEHMM_WARPS_PER_CTA * 32;
    const bool tr = (flags & TEHMM_BWD_TRANS) != 0;
#define BWD_GO(R, TR) backward_kernel<T, NS, R, TR><<<grid, th, 0, st>>>(m, b, flags, blin, alpha, ratios, post, map_states, map_part, xi_part, xdiag_part, gamma0, start_vec, end_vec, bad, mode)
    if (ratios) { if (tr) BWD_GO(true, true); else BWD_GO(true, false); }
    else { if (tr) BWD_GO(false, true); else BWD_GO(false, false); }
#undef BWD_GO
    return cudaGetLastError();
}

cudaError_t tehmm_launch_backward(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                                  int prec, int flags, const void *blin, const void *alpha,
                                  const double *ratios, void *post, uint8_t *map_states,
                                  double *map_part, void *xi_part, void *xdiag_part, void *gamma0,
                                  void *start_vec, void *end_vec, const int *bad, int mode, int grid)
{
    if (prec == TEHMM_F32) {
        if (m.NS == 1) return launch_bwd<float, 1>(st, m, b, flags, (const float *)blin, (const float *)alpha, ratios, (float *)post, map_states, map_part, (float *)xi_part, (float *)xdiag_part, (float *)gamma0, (float *)start_vec, (float *)end_vec, bad, mode, grid);
        return launch_bwd<float, 2>(st, m, b, flags, (const float *)blin, (const float *)alpha, ratios, (float *)post, map_states, map_part, (float *)xi_part, (float *)xdiag_part, (float *)gamma0, (float *)start_vec, (float *)end_vec, bad, mode, grid);
    }
    if (m.NS == 1) return launch_bwd<double, 1>(st, m, b, flags, (const double *)blin, (const double *)alpha, ratios, (double *)post, map_states, map_part, (double *)xi_part, (double *)xdiag_part, (double *)gamma0, (double *)start_vec, (double *)end_vec, bad, mode, grid);
    return launch_bwd<double, 2>(st, m, b, flags, (const double *)blin, (const double *)alpha, ratios, (double *)post, map_states, map_part, (double *)xi_part, (double *)xdiag_part, (double *)gamma0, (double *)start_vec, (double *)end_vec, bad, mode, grid);
}

cudaError_t tehmm_launch_trans_reduce(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                                      int prec, const void *xi_part, const void *xdiag_part,
                                      const void *gamma0, double *start_trans)
{
    int cells = m.N * m.N;
    int grid = (cells + 63) / 64;
    if (prec == TEHMM_F32)
        trans_reduce_kernel<float><<<grid, 64, 0, st>>>(m, b, (const float *)xi_part, (const float *)xdiag_part, (const float *)gamma0, start_trans);
    else
        trans_reduce_kernel<double><<<grid, 64, 0, st>>>(m, b, (const double *)xi_part, (const double *)xdiag_part, (const double *)gamma0, start_trans);
    return cudaGetLastError();
}

cudaError_t tehmm_launch_map_reduce(cudaStream_t st, const TehmmBatchDev &b, const double *map_part,
                                    double *map_score)
{
    int warps = 4;
    int grid = (int)((b.nseq + warps - 1) / warps);
    map_reduce_kernel<<<grid, warps * 32, 0, st>>>(b, map_part, map_score);
    return cudaGetLastError();
}
