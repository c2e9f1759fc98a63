// Chunked parallel-in-time forward pass in scaled-probability space
// (hmm.py:678-713 -> _hmm.pyx:120-158).
//
// The time axis of every sequence is cut into chunks; one warp per chunk.
// A chunk that does not begin its sequence needs alpha at t0-1, which the
// serial recursion would only know after all earlier chunks.  SPECULATE: run
// `warmup` extra steps from a uniform vector before t0 (the forward filter
// forgets its initial condition geometrically); VERIFY: compare the speculated
// vector with the true end vector of the previous chunk; REPAIR: re-run only
// the chunks that disagree, from the true vector, until none disagree.
//
// Scaling is by exact powers of two (canonicalise), so a chunk's output does
// not depend on the scaling history, only on its (canonical) start vector.
//
// Per step and warp: 1 STS + NP/4 LDS.128 (broadcast) + NP*NS FFMA + 1 REDUX +
// 1 coalesced load of b_t (N values) + 1 coalesced store of alpha_t.
// Algorithmic HBM bytes per step: 4N read + 4N written (fp32).
#include "scan.cuh"

#define FWD_U 4   // register prefetch depth (time steps)

template <typename T, int NS, bool RATIO>
__global__ void __launch_bounds__(TEHMM_WARPS_PER_CTA * 32, (sizeof(T) == 4 && NS == 1) ? 3 : 1)
forward_kernel(TehmmModelDev m, TehmmBatchDev b, const T *__restrict__ blin,
               const double *__restrict__ rowmax, const double *__restrict__ ratios,
               T *__restrict__ alpha, T *__restrict__ start_vec, T *__restrict__ end_vec,
               double *__restrict__ cscale, const int *__restrict__ bad, int mode)
{
    constexpr int NP = 32 * NS;
    __shared__ __align__(16) T xs_all[TEHMM_WARPS_PER_CTA][2][NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T(*xs)[NP] = xs_all[warp];
    const int N = m.N;

    // column j of the transition matrix for each owned state (zero beyond N, so
    // lanes without a state compute exact zeros and need no predicates)
    MatSlice<T, NS> A;
    T pi[NS];
    double dg[NS];
    unsigned jc[NS];      // clamped column for loads
    bool own[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int j = lane + 32 * s;
#pragma unroll
        for (int i = 0; i < NP; ++i) A.set(s, i, (T)m.lin_trans[(int64_t)i * NP + j]);
        pi[s] = (T)m.lin_start[j];
        dg[s] = RATIO ? m.cut_trans[(int64_t)j * NP + j] : 0.0;
        own[s] = j < N;
        jc[s] = (unsigned)min(j, N - 1);
    }

    for (int64_t ci = (int64_t)blockIdx.x * TEHMM_WARPS_PER_CTA + warp; ci < b.nchunks;
         ci += (int64_t)gridDim.x * TEHMM_WARPS_PER_CTA) {
        if (mode == 1 && !bad[ci]) continue;
        const TehmmChunk ch = b.chunks[ci];
        T x[NS];
        int esum = 0;
        double lsum = 0.0;
        int buf = 0;

        // where this warp starts reading: the warm-up start (speculative pass) or t0
        int64_t tw = ch.t0;
        bool from_start = ch.t0 == ch.s0;
        if (mode == 0 && ch.t0 > ch.s0) {
            tw = ch.t0 - b.warmup;
            if (tw <= ch.s0) { tw = ch.s0; from_start = true; }
        }
        // everything below addresses rows relative to tw with 32-bit offsets
        const T *__restrict__ bb = blin + tw * m.LD;
        T *__restrict__ aa = alpha ? alpha + tw * m.LD : nullptr;
        const double *__restrict__ rr = RATIO ? ratios + tw : nullptr;

        // one recursion step: x <- scaled( (x A) .* b_t [.* g_t] ); returns the exponent
        auto step = [&](unsigned row, const T (&bt)[NS], bool first) -> int {
            T raw[NS];
            if (first) {
#pragma unroll
                for (int s = 0; s < NS; ++s) raw[s] = pi[s] * bt[s];
            } else {
#pragma unroll
                for (int s = 0; s < NS; ++s) xs[buf][lane + 32 * s] = x[s];
                __syncwarp();
                T y[NS];
                matvec_sum<NS>(xs[buf], A, y);
                buf ^= 1;
#pragma unroll
                for (int s = 0; s < NS; ++s) raw[s] = y[s] * bt[s];
            }
            if (RATIO) {
                const double r = rr[row];
                if (r > 1.0) {
                    double lg[NS], mg = -INFINITY;
#pragma unroll
                    for (int s = 0; s < NS; ++s) { lg[s] = dg[s] * (r - 1.0); mg = fmax(mg, lg[s]); }
                    mg = warp_max(mg);
                    if (mg > -INFINITY) {
#pragma unroll
                        for (int s = 0; s < NS; ++s) raw[s] *= (T)exp(lg[s] - mg);
                        lsum += mg;
                    } else {
#pragma unroll
                        for (int s = 0; s < NS; ++s) raw[s] = (T)0;
                    }
                }
            }
            const int e = canonicalise<T, NS>(raw);
#pragma unroll
            for (int s = 0; s < NS; ++s) x[s] = raw[s];
            return e;
        };
        // running pointers of this lane's element in the current row (bumped by N per row)
        const T *__restrict__ bp[NS];
        T *__restrict__ ap[NS];
        bool wr[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            bp[s] = bb + jc[s];
            ap[s] = aa + lane + 32 * s;
            wr[s] = lane + 32 * s < m.LD && aa != nullptr;   // padding columns get the exact zeros
        }
        const unsigned Nu = (unsigned)m.LD;
        // b of the row `ahead` rows after the current one
        auto load_b = [&](unsigned ahead, T (&bt)[NS]) {
#pragma unroll
            for (int s = 0; s < NS; ++s) bt[s] = bp[s][ahead * Nu];
        };
        auto store_a = [&]() {
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (wr[s]) *ap[s] = x[s];
        };
        auto advance = [&](unsigned rows) {
#pragma unroll
            for (int s = 0; s < NS; ++s) { bp[s] += rows * Nu; ap[s] += rows * Nu; }
        };

        unsigned row = 0;                                   // current row, relative to tw
        const unsigned row0 = (unsigned)(ch.t0 - tw);       // first row with output
        const unsigned row1 = (unsigned)(ch.t1 - tw);       // one past the last row
        if (mode == 1) {
#pragma unroll
            for (int s = 0; s < NS; ++s) x[s] = start_vec[ci * NP + lane + 32 * s];
        } else if (from_start) {
            T bt[NS];
            load_b(0, bt);
            const int e = step(0, bt, true);
            if (row0 == 0) { esum += e; store_a(); }
            row = 1;
            advance(1);
        } else {
#pragma unroll
            for (int s = 0; s < NS; ++s) x[s] = own[s] ? (T)1 : (T)0;
        }
        if (mode == 0 && row0 > 0) {
            // speculative warm-up: no output, scale discarded
            for (; row < row0; ++row) {
                T bt[NS];
                load_b(0, bt);
                step(row, bt, false);
                advance(1);
            }
            lsum = 0.0;
#pragma unroll
            for (int s = 0; s < NS; ++s) start_vec[ci * NP + lane + 32 * s] = x[s];
        }

        // main loop: groups of FWD_U steps, the next group's b rows already in flight
        T bn[FWD_U][NS];
        if (row + FWD_U <= row1) {
#pragma unroll
            for (int u = 0; u < FWD_U; ++u) load_b(u, bn[u]);
        }
        while (row + 2 * FWD_U <= row1) {
            T bc[FWD_U][NS];
#pragma unroll
            for (int u = 0; u < FWD_U; ++u) {
#pragma unroll
                for (int s = 0; s < NS; ++s) bc[u][s] = bn[u][s];
            }
#pragma unroll
            for (int u = 0; u < FWD_U; ++u) load_b(FWD_U + u, bn[u]);
#pragma unroll
            for (int u = 0; u < FWD_U; ++u) {
                esum += step(row + u, bc[u], false);
                store_a();
                advance(1);
            }
            row += FWD_U;
        }
        if (row + FWD_U <= row1) {
#pragma unroll
            for (int u = 0; u < FWD_U; ++u) {
                esum += step(row + u, bn[u], false);
                store_a();
                advance(1);
            }
            row += FWD_U;
        }
        for (; row < row1; ++row) {
            T bt[NS];
            load_b(0, bt);
            esum += step(row, bt, false);
            store_a();
            advance(1);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) end_vec[ci * NP + lane + 32 * s] = x[s];

        // log of everything taken out of the chunk: exponents, ratio maxima, row maxima
        double ms = 0.0;
        for (int64_t tt = ch.t0 + lane; tt < ch.t1; tt += 32) ms += rowmax[tt];
        ms = warp_sum(ms);
        if (lane == 0) cscale[ci] = (double)esum * 0.6931471805599453094 + lsum + ms;
    }
}

// bad[c] = speculated start vector of chunk c differs from the neighbour's end
// vector by more than tol.  dir=+1: neighbour is c-1 (forward, Viterbi);
// dir=-1: neighbour is c+1 (backward).
// Linear-space vectors (forward / backward) are defined only up to a scale
// factor, and what the recursion contracts is Hilbert's projective metric, so
// they are compared by the spread of the component-wise ratios:
//     max_j(a_j/t_j) / min_j(a_j/t_j) - 1 <= tol
// (components that are negligible in both vectors are skipped).  A small
// component matters as much as a large one here because the other pass
// (beta for alpha, alpha for beta) may weight it up in the posterior.
// Log-space vectors (Viterbi) are compared by the spread of the differences.
// A bad chunk gets the true vector copied into its start slot for the repair.
// One warp per chunk, lane = state (and state + 32 for 33..64 states): the vectors are read as
// coalesced rows and min / max / sums are warp reductions (with one thread per chunk walking 32
// float64 divisions the kernel took 15 us per pass, four passes per sweep).
__device__ __forceinline__ double vwarp_fmax(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(TEHMM_FULL, v, o));   // fmax / fmin skip NaNs
    return v;
}
__device__ __forceinline__ double vwarp_fmin(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(TEHMM_FULL, v, o));
    return v;
}
template <typename T>
__global__ void verify_kernel(TehmmBatchDev b, int NP, T *start_vec, const T *end_vec,
                              double tol, int dir, int linear, int *bad, int *nbad,
                              double *logkappa)
{
    const int lane = threadIdx.x & 31;
    const int64_t ci = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ci >= b.nchunks) return;                       // the whole warp leaves together
    const TehmmChunk ch = b.chunks[ci];
    const bool has = dir > 0 ? (ch.t0 > ch.s0) : (ch.t1 < ch.s1);
    int flag = 0;
    double lk = 0.0;
    if (has) {
        const T *tv = end_vec + (ci - dir) * NP;
        T *sv = start_vec + ci * NP;
        const int nj = NP >> 5;                        // 1 or 2 components per lane
        double a[2] = {0.0, 0.0}, t[2] = {0.0, 0.0};
        for (int u = 0; u < nj; ++u) { a[u] = (double)sv[lane + 32 * u]; t[u] = (double)tv[lane + 32 * u]; }
        double lo = INFINITY, hi = -INFINITY;
        int mismatch = 0;
        if (linear) {
            double ma = 0.0, mt = 0.0;
            for (int u = 0; u < nj; ++u) { ma = fmax(ma, a[u]); mt = fmax(mt, t[u]); }
            ma = vwarp_fmax(ma);
            mt = vwarp_fmax(mt);
            const double floor_rel = sizeof(T) == 4 ? 1e-30 : 1e-250;
            for (int u = 0; u < nj; ++u) {
                const bool az = !(a[u] > floor_rel * ma), tz = !(t[u] > floor_rel * mt);
                if (az && tz) continue;
                if (az != tz) { mismatch = 1; continue; }
                const double r = a[u] / t[u];
                lo = fmin(lo, r); hi = fmax(hi, r);
            }
            lo = vwarp_fmin(lo);
            hi = vwarp_fmax(hi);
            flag = __any_sync(TEHMM_FULL, mismatch) ? 1 : 0;
            if (!flag && hi > -INFINITY && !(hi / lo - 1.0 <= tol)) flag = 1;   // NaN counts as bad
        } else {
            for (int u = 0; u < nj; ++u) {
                const bool az = !(a[u] > -1e30), tz = !(t[u] > -1e30);
                if (az && tz) continue;
                if (az != tz) { mismatch = 1; continue; }
                const double d = a[u] - t[u];
                lo = fmin(lo, d); hi = fmax(hi, d);
            }
            lo = vwarp_fmin(lo);
            hi = vwarp_fmax(hi);
            flag = __any_sync(TEHMM_FULL, mismatch) ? 1 : 0;
            if (!flag && hi > -INFINITY && !(hi - lo <= tol)) flag = 1;
        }
        if (flag) {
            for (int u = 0; u < nj; ++u) sv[lane + 32 * u] = tv[lane + 32 * u];
        } else if (linear && logkappa) {
            // scale of the speculated vector relative to the true one (see forward_logprob_kernel)
            double ss = 0.0, se = 0.0;
            for (int u = 0; u < nj; ++u) { ss += a[u]; se += t[u]; }
            ss = warp_sum(ss);
            se = warp_sum(se);
            lk = log(ss / se);
        }
    }
    if (lane == 0) {
        bad[ci] = flag;
        if (logkappa) logkappa[ci] = lk;
        if (flag) atomicAdd(nbad, 1);
    }
}

// logprob[seq] = sum of chunk scales + log(sum_j alpha_hat[T-1][j])
//                - sum over chunk boundaries of log(kappa_c),
// kappa_c = sum(start_vec[c]) / sum(end_vec[c-1]) (written by verify_kernel): a
// chunk measures its scale relative to its own (speculated) start vector,
// which points in the same direction as the previous chunk's end vector but
// differs from it by an arbitrary factor in (1/2, 2).
template <typename T>
__global__ void forward_logprob_kernel(TehmmBatchDev b, int NP, const T *__restrict__ end_vec,
                                       const double *__restrict__ cscale,
                                       const double *__restrict__ logkappa,
                                       double *__restrict__ logprob)
{
    // one block per sequence
    const int64_t s = blockIdx.x;
    const int64_t c0 = b.seq_chunk0[s], c1 = b.seq_chunk0[s + 1];
    if (c1 <= c0) { if (threadIdx.x == 0) logprob[s] = 0.0; return; }
    double acc = 0.0;
    for (int64_t c = c0 + threadIdx.x; c < c1; c += blockDim.x) acc += cscale[c] - logkappa[c];
    acc = block_sum(acc);
    double tail = threadIdx.x < NP ? (double)end_vec[(c1 - 1) * NP + threadIdx.x] : 0.0;
    tail = block_sum(tail);
    if (threadIdx.x == 0) logprob[s] = acc + log(tail);
}

template <typename T, int NS>
static cudaError_t launch_fwd(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                              const T *blin, const double *rowmax, const double *ratios, T *alpha,
                              T *start_vec, T *end_vec, double *cscale, const int *bad, int mode,
                              int grid)
{
    if (ratios)
        forward_kernel<T, NS, true><<<grid, TEHMM_WARPS_PER_CTA * 32, 0, st>>>(m, b, blin, rowmax, ratios, alpha, start_vec, end_vec, cscale, bad, mode);
    else
        forward_kernel<T, NS, false><<<grid, TEHMM_WARPS_PER_CTA * 32, 0, st>>>(m, b, blin, rowmax, ratios, alpha, start_vec, end_vec, cscale, bad, mode);
    return cudaGetLastError();
}

cudaError_t tehmm_launch_forward(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                                 int prec, const void *blin, const double *rowmax,
                                 const double *ratios, void *alpha, void *start_vec, void *end_vec,
                                 double *cscale, const int *bad, int mode, int grid)
{
    if (prec == TEHMM_F32) {
        if (m.NS == 1) return launch_fwd<float, 1>(st, m, b, (const float *)blin, rowmax, ratios, (float *)alpha, (float *)start_vec, (float *)end_vec, cscale, bad, mode, grid);
        return launch_fwd<float, 2>(st, m, b, (const float *)blin, rowmax, ratios, (float *)alpha, (float *)start_vec, (float *)end_vec, cscale, bad, mode, grid);
    }
    if (m.NS == 1) return launch_fwd<double, 1>(st, m, b, (const double *)blin, rowmax, ratios, (double *)alpha, (double *)start_vec, (double *)end_vec, cscale, bad, mode, grid);
    return launch_fwd<double, 2>(st, m, b, (const double *)blin, rowmax, ratios, (double *)alpha, (double *)start_vec, (double *)end_vec, cscale, bad, mode, grid);
}

cudaError_t tehmm_launch_verify(cudaStream_t st, const TehmmBatchDev &b, int prec, int NP,
                                void *start_vec, const void *end_vec, double tol, int dir,
                                int linear, int *bad, int *nbad, double *logkappa)
{
    cudaError_t e = cudaMemsetAsync(nbad, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    int grid = (int)((b.nchunks + 3) / 4);             // one warp per chunk, four warps per block
    if (prec == TEHMM_F32)
        verify_kernel<float><<<grid, 128, 0, st>>>(b, NP, (float *)start_vec, (const float *)end_vec, tol, dir, linear, bad, nbad, logkappa);
    else
        verify_kernel<double><<<grid, 128, 0, st>>>(b, NP, (double *)start_vec, (const double *)end_vec, tol, dir, linear, bad, nbad, logkappa);
    return cudaGetLastError();
}

cudaError_t tehmm_launch_forward_logprob(cudaStream_t st, const TehmmBatchDev &b, int prec, int NP,
                                         const void *end_vec, const double *cscale,
                                         const double *logkappa, double *logprob)
{
    // one block per sequence; its chunks are read in a strided loop of dependent-latency loads, so long
    // sequences get the widest block (24 us -> a few at 18 944 chunks)
    const int64_t cps = b.nchunks / b.nseq;
    const int th = cps >= 2048 ? 1024 : cps >= 256 ? 256 : 64;
    if (prec == TEHMM_F32)
        forward_logprob_kernel<float><<<(int)b.nseq, th, 0, st>>>(b, NP, (const float *)end_vec, cscale, logkappa, logprob);
    else
        forward_logprob_kernel<double><<<(int)b.nseq, th, 0, st>>>(b, NP, (const double *)end_vec, cscale, logkappa, logprob);
    return cudaGetLastError();
}
