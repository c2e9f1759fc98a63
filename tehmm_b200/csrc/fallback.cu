// Bounded-cost exact resolution of chunk boundaries (the "chunked parallel-in-time semiring scan"
// of BASELINE.json's north_star, used where speculation does not pay).
//
// forward.cu / backward.cu / viterbi.cu speculate a chunk's boundary vector with a short warm-up
// and repair the chunks whose speculation failed, one pass per link of a chain of bad chunks.
// That is the right trade when filters forget quickly (0 repairs on the bench input), and the
// wrong one on slow-mixing input -- long all-missing stretches under a sticky transition matrix:
// there the truth travels ONE chunk per pass (probe: 91 passes for eight gaps of <= 20 000 steps,
// tools/probe_hard.py).  After a few ordinary passes the driver (api.cu, resolve_flagged) therefore
// switches to the exact scheme:
//   1. for every chunk still flagged, its N x N transfer operator over the chunk's rows, one warp
//      per (chunk, basis state) -- this kernel.  Sum-product operators in scaled space with the
//      exponent taken out returned separately (rows of one operator carry different scales),
//      (max,+) operators with the normaliser returned likewise;
//   2. a chain in float64 over each run of flagged chunks (N^2 work per chunk, one warp per run);
//   3. ONE re-run of the flagged chunks from the chained vectors by the ordinary kernels.
// Cost: N basis walks per FLAGGED chunk (one extra wave for up to ~120 flagged chunks of the fine
// partition at 30 states, proportional beyond), instead of one pass per link.
//
// The reference has no counterpart: its recursions are serial (_hmm.pyx:120-259).
#include "scan.cuh"

// KIND 0: forward, sum-product:   x <- canon((x A) .* b_t),            t = t0 .. t1-1, x0 = e_i
// KIND 1: backward, sum-product:  u <- canon(A (b_{t+1} .* u)),         t = t1-1 .. t0, u(t1) = e_i
// KIND 2: forward, (max,+):       d <- norm(max_i(d_i + logA_ij) + e_t), t = t0 .. t1-1, d0 = 0 at i, -inf elsewhere
template <typename T, int NS, int KIND>
__global__ void __launch_bounds__(TEHMM_WARPS_PER_CTA * 32)
transfer_op_kernel(TehmmModelDev m, const TehmmChunk *__restrict__ chunks, const int64_t *__restrict__ flagged,
                   int64_t nflag, const T *__restrict__ lat, T *__restrict__ op_end, double *__restrict__ op_scale)
{
    constexpr int NP = 32 * NS;
    __shared__ __align__(16) T xs_all[TEHMM_WARPS_PER_CTA][2][NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T(*xs)[NP] = xs_all[warp];
    const int N = m.N;
    const unsigned Nu = (unsigned)m.LD;
    MatSlice<T, NS> A;
    unsigned jc[NS];
    bool own[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int j = lane + 32 * s;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (KIND == 0) A.set(s, i, (T)m.lin_trans[(int64_t)i * NP + j]);        // column j
            else if (KIND == 1) A.set(s, i, (T)m.lin_trans[(int64_t)j * NP + i]);   // row j
            else A.set(s, i, (T)m.cut_trans[(int64_t)i * NP + j]);                  // column j of log A
        }
        own[s] = j < N;
        jc[s] = (unsigned)min(j, N - 1);
    }
    const int64_t nv = nflag * N;
    for (int64_t v = (int64_t)blockIdx.x * TEHMM_WARPS_PER_CTA + warp; v < nv;
         v += (int64_t)gridDim.x * TEHMM_WARPS_PER_CTA) {
        const int64_t k = v / N;
        const int basis = (int)(v - k * N);
        const TehmmChunk ch = chunks[flagged[k]];
        const T *__restrict__ bb = lat + ch.t0 * m.LD;
        const unsigned nrows = (unsigned)(ch.t1 - ch.t0);
        T x[NS];
        int buf = 0;
        double acc = 0.0;          // log of everything taken out
        if (KIND == 2) {
#pragma unroll
            for (int s = 0; s < NS; ++s) x[s] = lane + 32 * s == basis ? (T)0 : (T)-INFINITY;
            for (unsigned r = 0; r < nrows; ++r) {
                T et[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) et[s] = own[s] ? bb[r * Nu + jc[s]] : (T)-INFINITY;
#pragma unroll
                for (int s = 0; s < NS; ++s) xs[buf][lane + 32 * s] = x[s];
                __syncwarp();
                T y[NS];
                matvec_maxval<NS>(xs[buf], A, y);
                buf ^= 1;
                T mx = (T)-INFINITY;
#pragma unroll
                for (int s = 0; s < NS; ++s) { x[s] = y[s] + et[s]; mx = x[s] > mx ? x[s] : mx; }
                mx = warp_max(mx);
                if (mx > (T)-INFINITY) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) x[s] -= mx;
                    acc += (double)mx;
                } else {
                    acc = -INFINITY;       // state `basis` cannot reach the end of the chunk
                }
            }
        } else {
            int esum = 0;
#pragma unroll
            for (int s = 0; s < NS; ++s) x[s] = lane + 32 * s == basis ? (T)1 : (T)0;
            for (unsigned q = 0; q < nrows; ++q) {
                // forward: row q; backward: computes beta at row nrows-1-q from b at the row after it
                const unsigned r = KIND == 0 ? q : nrows - q;
                T bt[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) bt[s] = own[s] ? bb[r * Nu + jc[s]] : (T)0;
                T y[NS];
                if (KIND == 0) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) xs[buf][lane + 32 * s] = x[s];
                    __syncwarp();
                    matvec_sum<NS>(xs[buf], A, y);
#pragma unroll
                    for (int s = 0; s < NS; ++s) y[s] *= bt[s];
                } else {
#pragma unroll
                    for (int s = 0; s < NS; ++s) xs[buf][lane + 32 * s] = bt[s] * x[s];
                    __syncwarp();
                    matvec_sum<NS>(xs[buf], A, y);
                }
                buf ^= 1;
                esum += canonicalise<T, NS>(y);
#pragma unroll
                for (int s = 0; s < NS; ++s) x[s] = y[s];
            }
            acc = (double)esum * 0.6931471805599453094;
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) op_end[v * NP + lane + 32 * s] = x[s];
        if (lane == 0) op_scale[v] = acc;
    }
}

// The chain over each run of consecutive flagged chunks, one warp per run (the links of a run are
// serial; runs are independent).  `list` holds the flagged chunks in chain order, run after run, and
// operator k belongs to list[k].  Link: the chunk takes `cur` as its boundary vector (written to
// start_vec) and hands on operator(cur), renormalised; a run starts from the stored end vector of its
// standing neighbour.  float64 throughout: the rows of one operator carry different scales.
template <typename T>
__global__ void chain_kernel(int N, int NP, int kind, int dir, const int64_t *__restrict__ list,
                             const int64_t *__restrict__ run_off, int64_t nruns, const T *__restrict__ op_end,
                             const double *__restrict__ op_scale, const T *__restrict__ ev, T *__restrict__ sv)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nruns) return;
    const int nu = NP >> 5;                      // 1 or 2 components per lane
    double cur[2] = {0.0, 0.0};
    for (int64_t k = run_off[r]; k < run_off[r + 1]; ++k) {
        const int64_t cidx = list[k];
        if (k == run_off[r])
            for (int u = 0; u < nu; ++u) cur[u] = (double)ev[(cidx - dir) * NP + lane + 32 * u];
        for (int u = 0; u < nu; ++u) sv[cidx * NP + lane + 32 * u] = (T)cur[u];
        const T *E = op_end + (int64_t)k * N * NP;
        const double *S = op_scale + (int64_t)k * N;
        double nx[2];
        if (kind == 2) {
            nx[0] = nx[1] = -INFINITY;
            for (int i = 0; i < N; ++i) {
                const double ci = __shfl_sync(TEHMM_FULL, cur[i >> 5], i & 31) + S[i];
                for (int u = 0; u < nu; ++u) {
                    const double v = ci + (double)E[(int64_t)i * NP + lane + 32 * u];
                    if (v > nx[u]) nx[u] = v;            // NaN (inf - inf) never wins
                }
            }
            double mx = -INFINITY;
            for (int u = 0; u < nu; ++u) { if (lane + 32 * u >= N) nx[u] = -INFINITY; mx = fmax(mx, nx[u]); }
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(TEHMM_FULL, mx, o));
            for (int u = 0; u < nu; ++u) cur[u] = mx > -INFINITY ? nx[u] - mx : nx[u];
        } else {
            double smax = -INFINITY;
            for (int u = 0; u < nu; ++u) {
                const int i = lane + 32 * u;
                if (i < N && cur[u] > 0.0) smax = fmax(smax, S[i]);
            }
            for (int o = 16; o > 0; o >>= 1) smax = fmax(smax, __shfl_xor_sync(TEHMM_FULL, smax, o));
            double w[2];
            for (int u = 0; u < nu; ++u) {
                const int i = lane + 32 * u;
                w[u] = (i < N && cur[u] > 0.0 && smax > -INFINITY) ? cur[u] * exp(S[i] - smax) : 0.0;
            }
            nx[0] = nx[1] = 0.0;
            for (int i = 0; i < N; ++i) {
                const double wi = __shfl_sync(TEHMM_FULL, w[i >> 5], i & 31);
                for (int u = 0; u < nu; ++u) nx[u] = fma(wi, (double)E[(int64_t)i * NP + lane + 32 * u], nx[u]);
            }
            double mx = 0.0;
            for (int u = 0; u < nu; ++u) { if (lane + 32 * u >= N) nx[u] = 0.0; mx = fmax(mx, nx[u]); }
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(TEHMM_FULL, mx, o));
            for (int u = 0; u < nu; ++u) cur[u] = mx > 0.0 ? nx[u] / mx : 0.0;
        }
    }
}

cudaError_t tehmm_launch_chain(cudaStream_t st, const TehmmModelDev &m, int prec, int kind, int dir, const int64_t *list,
                               const int64_t *run_off, int64_t nruns, const void *op_end, const double *op_scale,
                               const void *ev, void *sv)
{
    const int grid = (int)((nruns + 3) / 4);
    if (prec == TEHMM_F32)
        chain_kernel<float><<<grid, 128, 0, st>>>(m.N, m.NP, kind, dir, list, run_off, nruns, (const float *)op_end, op_scale, (const float *)ev, (float *)sv);
    else
        chain_kernel<double><<<grid, 128, 0, st>>>(m.N, m.NP, kind, dir, list, run_off, nruns, (const double *)op_end, op_scale, (const double *)ev, (double *)sv);
    return cudaGetLastError();
}

template <typename T, int NS>
static cudaError_t launch_ops(cudaStream_t st, const TehmmModelDev &m, const TehmmChunk *chunks, const int64_t *flagged,
                              int64_t nflag, int kind, const T *lat, T *op_end, double *op_scale, int sms)
{
    const int64_t need = (nflag * m.N + TEHMM_WARPS_PER_CTA - 1) / TEHMM_WARPS_PER_CTA;
    const int grid = (int)(need < (int64_t)sms * 8 ? (need < 1 ? 1 : need) : (int64_t)sms * 8);
    const int th = TEHMM_WARPS_PER_CTA * 32;
    if (kind == 0) transfer_op_kernel<T, NS, 0><<<grid, th, 0, st>>>(m, chunks, flagged, nflag, lat, op_end, op_scale);
    else if (kind == 1) transfer_op_kernel<T, NS, 1><<<grid, th, 0, st>>>(m, chunks, flagged, nflag, lat, op_end, op_scale);
    else transfer_op_kernel<T, NS, 2><<<grid, th, 0, st>>>(m, chunks, flagged, nflag, lat, op_end, op_scale);
    return cudaGetLastError();
}

// kind: 0 forward (lat = blin), 1 backward (lat = blin), 2 Viterbi (lat = elog)
cudaError_t tehmm_launch_transfer_ops(cudaStream_t st, const TehmmModelDev &m, const TehmmChunk *chunks,
                                      const int64_t *flagged, int64_t nflag, int prec, int kind, const void *lat,
                                      void *op_end, double *op_scale, int sms)
{
    if (prec == TEHMM_F32) {
        if (m.NS == 1) return launch_ops<float, 1>(st, m, chunks, flagged, nflag, kind, (const float *)lat, (float *)op_end, op_scale, sms);
        return launch_ops<float, 2>(st, m, chunks, flagged, nflag, kind, (const float *)lat, (float *)op_end, op_scale, sms);
    }
    if (m.NS == 1) return launch_ops<double, 1>(st, m, chunks, flagged, nflag, kind, (const double *)lat, (double *)op_end, op_scale, sms);
    return launch_ops<double, 2>(st, m, chunks, flagged, nflag, kind, (const double *)lat, (double *)op_end, op_scale, sms);
}
