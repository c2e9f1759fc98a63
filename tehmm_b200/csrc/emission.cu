// Batched emission gather-and-sum (emission.py:179-198 -> _emission.pyx:50-80).
//
// HBM-bound: reads K bytes of symbols per time step and writes N values (twice
// when both the log-space and the linear copy are requested) + one float64 row
// maximum.  The per-track tables are staged once per CTA in shared memory,
// transposed to [symbol-row][state] so that the 32 lanes of a warp (= states)
// read consecutive words (conflict-free).  One warp per time step, 32 steps of
// symbols staged per warp iteration with a coalesced load.
#include "scan.cuh"

#define EM_WARPS 16
#define EM_ROWS 16   // time steps staged per warp iteration

// MODE 0: write normalised log (elog) / linear (blin) copies in type T + rowmax
// MODE 1: write the reference-layout float64 frame only
// Staging turns each symbol into the element offset of its table row
// ((tab_off[k] + symbol) * N), so the inner loop is one broadcast LDS of the
// offset and one conflict-free LDS.64 of the table entry per (row, track).
// A symbol outside the compact table is staged as -(1 + k*S + symbol) and read
// from the dense table exactly as the reference indexes it (rare, per-row vote).
// Two rows are in flight per iteration: the sum over tracks must stay
// sequential (bit-exact against the reference), so the ILP comes from rows.
template <typename T, int MODE, int NS>
__device__ __forceinline__ void em_finish_row(const TehmmModelDev &m, int64_t t, double (&v)[NS],
                                              const double *__restrict__ ratios, T *elog, T *blin,
                                              double *rowmax, double *frame, int *seq_flag,
                                              const int64_t *seq_off, int64_t nseq, int lane)
{
    const int N = m.N;
    double vmax = -INFINITY;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        v[s] *= m.normalize;
        if (ratios) v[s] *= ratios[t];
        if (lane + 32 * s < N) vmax = fmax(vmax, v[s]); else v[s] = -INFINITY;
    }
    // row maximum: float REDUX first; exact double fallback when it overflows fp32
    float mf = warp_max_any((float)vmax);
    double M = (double)mf;
    if (!(mf > (float)TEHMM_MINDBL)) {
        M = warp_max(vmax);
        // candidate for the "no state can emit" quirk (_emission.pyx:73-80)
        if (!(M > TEHMM_MINDBL) && seq_flag && lane == 0) {
            int64_t lo = 0, hi = nseq;   // sequence of row t
            while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (seq_off[mid] <= t) lo = mid; else hi = mid; }
            seq_flag[lo] = 1;
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (lane + 32 * s < N) frame[t * N + lane + 32 * s] = v[s];
        return;
    }
    if (lane == 0) rowmax[t] = M;
    const int64_t o = t * m.LD + lane;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (lane + 32 * s >= N && lane + 32 * s < m.LD) {        // padding columns: zeros
            if (elog) elog[o + 32 * s] = (T)0;
            if (blin) blin[o + 32 * s] = (T)0;
        }
        if (lane + 32 * s < N) {
            const double d = (M > -INFINITY) ? v[s] - M : 0.0;
            if (sizeof(T) == 4) {
                // d <= 0 and the entries that matter are within a few units of 0, where
                // ex2.approx(d*log2e) is good to ~1e-7 relative
                const float df = (float)d;
                if (elog) elog[o + 32 * s] = (T)df;
                if (blin) blin[o + 32 * s] = (T)__expf(df);
            } else {
                if (elog) elog[o + 32 * s] = (T)d;
                if (blin) blin[o + 32 * s] = (T)exp(d);
            }
        }
    }
}

template <typename T, typename OBS, int MODE, int NS, bool SMEM>
__global__ void __launch_bounds__(EM_WARPS * 32, 2)
emission_kernel(TehmmModelDev m, const OBS *__restrict__ obs, int64_t total,
                const double *__restrict__ ratios, T *__restrict__ elog, T *__restrict__ blin,
                double *__restrict__ rowmax, double *__restrict__ frame,
                int *__restrict__ seq_flag, const int64_t *__restrict__ seq_off, int64_t nseq)
{
    extern __shared__ __align__(16) unsigned char em_smem[];
    // layout: [K] offs | [K] nsyms | per-warp offset staging | table
    int32_t *offs = reinterpret_cast<int32_t *>(em_smem);
    int32_t *nsyms = offs + m.K;
    int *stage_all = reinterpret_cast<int *>(nsyms + m.K);
    size_t head = ((size_t)(2 * m.K + EM_WARPS * EM_ROWS * m.K) * 4 + 15) & ~(size_t)15;
    double *tab_s = reinterpret_cast<double *>(em_smem + head);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = m.K, N = m.N;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        offs[k] = m.tab_off[k];
        nsyms[k] = m.track_nsym[k];
    }
    if (SMEM) {
        int64_t n = (int64_t)m.tab_rows * N;
        for (int64_t e = threadIdx.x; e < n; e += blockDim.x) tab_s[e] = m.table_t[e];
    }
    __syncthreads();
    const double *__restrict__ tab_g = m.table_t;

    auto entry = [&](int o, int j) -> double {
        if (SMEM) return tab_s[o + j];
        return tab_g[o + j];
    };
    auto slow_row = [&](const int *so, double (&v)[NS]) {
        for (int k = 0; k < K; ++k) {
            const int o = so[k];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const int j = lane + 32 * s;
                if (j < N) {
                    if (o >= 0) v[s] += entry(o, j);
                    else {
                        const int code = -o - 1;
                        v[s] += m.table[((int64_t)(code / m.S) * N + j) * m.S + code % m.S];
                    }
                }
            }
        }
    };

    int *stage = stage_all + warp * EM_ROWS * K;
    const int64_t nblocks = (total + EM_ROWS - 1) / EM_ROWS;
    for (int64_t blk = (int64_t)blockIdx.x * EM_WARPS + warp; blk < nblocks;
         blk += (int64_t)gridDim.x * EM_WARPS) {
        const int64_t tb = blk * EM_ROWS;
        const int rows = (int)min((int64_t)EM_ROWS, total - tb);
        __syncwarp();
        bool neg = false;
        for (int e = lane; e < rows * K; e += 32) {
            const int sym = (int)obs[tb * K + e];
            const int k = e % K;
            const int o = sym < nsyms[k] ? (offs[k] + sym) * N : -(1 + k * m.S + sym);
            neg |= o < 0;
            stage[e] = o;
        }
        const bool any_slow = __any_sync(TEHMM_FULL, neg);
        __syncwarp();
        // lanes beyond N read a valid (clamped) column and are discarded later
        int jc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) jc[s] = min(lane + 32 * s, N - 1);
        int r = 0;
        if (!any_slow) {
            for (; r + 1 < rows; r += 2) {
                const int *so0 = stage + r * K, *so1 = so0 + K;
                double v0[NS], v1[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) { v0[s] = 0.0; v1[s] = 0.0; }
#pragma unroll 5
                for (int k = 0; k < K; ++k) {
                    const int o0 = so0[k], o1 = so1[k];
#pragma unroll
                    for (int s = 0; s < NS; ++s) { v0[s] += entry(o0, jc[s]); v1[s] += entry(o1, jc[s]); }
                }
                em_finish_row<T, MODE, NS>(m, tb + r, v0, ratios, elog, blin, rowmax, frame, seq_flag, seq_off, nseq, lane);
                em_finish_row<T, MODE, NS>(m, tb + r + 1, v1, ratios, elog, blin, rowmax, frame, seq_flag, seq_off, nseq, lane);
            }
        }
        for (; r < rows; ++r) {
            double v[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) v[s] = 0.0;
            slow_row(stage + r * K, v);
            em_finish_row<T, MODE, NS>(m, tb + r, v, ratios, elog, blin, rowmax, frame, seq_flag, seq_off, nseq, lane);
        }
    }
}

// _emission.pyx:59,73-80: the running maximum is never reset, so rows are
// zeroed only while no earlier row of the sequence had a value > -1e20.
// One warp per flagged sequence; almost never runs.
template <typename T>
__global__ void emission_fix_kernel(int N, int LD, const int *__restrict__ seq_flag,
                                    const int64_t *__restrict__ seq_off, int64_t nseq,
                                    T *elog, T *blin, double *rowmax, double *frame)
{
    int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (s >= nseq || !seq_flag[s]) return;
    for (int64_t t = seq_off[s]; t < seq_off[s + 1]; ++t) {
        bool infeasible;
        if (frame) {
            double mx = -INFINITY;
            for (int j = lane; j < N; j += 32) mx = fmax(mx, frame[t * N + j]);
            mx = warp_max(mx);
            infeasible = !(mx > TEHMM_MINDBL);
        } else {
            infeasible = !(rowmax[t] > TEHMM_MINDBL);
        }
        if (!infeasible) break;
        for (int j = lane; j < N; j += 32) {
            if (frame) frame[t * N + j] = 0.0;
            if (elog) elog[t * LD + j] = (T)0;
            if (blin) blin[t * LD + j] = (T)1;
        }
        if (rowmax && lane == 0) rowmax[t] = 0.0;
        __syncwarp();
    }
}

static size_t em_smem_bytes(const TehmmModelDev &m)
{
    size_t head = ((size_t)(2 * m.K + EM_WARPS * EM_ROWS * m.K) * 4 + 15) & ~(size_t)15;
    return head + (m.table_in_smem ? (size_t)m.tab_rows * m.N * sizeof(double) : 0);
}

size_t tehmm_emission_table_budget(int K)
{
    size_t head = ((size_t)(2 * K + EM_WARPS * EM_ROWS * K) * 4 + 15) & ~(size_t)15;
    return (size_t)220 * 1024 - head;
}

template <typename T, typename OBS, int MODE, int NS>
static cudaError_t launch_em_ns(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                                const double *ratios, T *elog, T *blin, double *rowmax,
                                double *frame, int *seq_flag, int sms)
{
    size_t smem = em_smem_bytes(m);
    auto kern = m.table_in_smem ? emission_kernel<T, OBS, MODE, NS, true> : emission_kernel<T, OBS, MODE, NS, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = smem > 112 * 1024 ? 1 : 2;
    int64_t need = ((b.total + EM_ROWS - 1) / EM_ROWS + EM_WARPS - 1) / EM_WARPS;
    if (need < 1) need = 1;
    int64_t cap = (int64_t)sms * per_sm;
    int grid = (int)(need < cap ? need : cap);
    kern<<<grid, EM_WARPS * 32, smem, st>>>(m, (const OBS *)b.obs, b.total, ratios, elog, blin,
                                            rowmax, frame, seq_flag, b.seq_off, b.nseq);
    return cudaGetLastError();
}

template <typename T, typename OBS, int MODE>
static cudaError_t launch_em(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                             const double *ratios, T *elog, T *blin, double *rowmax,
                             double *frame, int *seq_flag, int sms)
{
    if (m.NS == 1) return launch_em_ns<T, OBS, MODE, 1>(st, m, b, ratios, elog, blin, rowmax, frame, seq_flag, sms);
    return launch_em_ns<T, OBS, MODE, 2>(st, m, b, ratios, elog, blin, rowmax, frame, seq_flag, sms);
}

template <typename T, int MODE>
static cudaError_t launch_em_obs(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                                 const double *ratios, T *elog, T *blin, double *rowmax,
                                 double *frame, int *seq_flag, int sms)
{
    if (b.obs_bytes == 1) return launch_em<T, uint8_t, MODE>(st, m, b, ratios, elog, blin, rowmax, frame, seq_flag, sms);
    if (b.obs_bytes == 2) return launch_em<T, uint16_t, MODE>(st, m, b, ratios, elog, blin, rowmax, frame, seq_flag, sms);
    return launch_em<T, int32_t, MODE>(st, m, b, ratios, elog, blin, rowmax, frame, seq_flag, sms);
}

// returns number of kernels launched, or -1 with *err set
int tehmm_launch_emission(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b, int prec,
                          const double *ratios, void *elog, void *blin, double *rowmax,
                          double *frame, int *seq_flag, int sms, cudaError_t *err)
{
    cudaError_t e = cudaMemsetAsync(seq_flag, 0, sizeof(int) * (size_t)b.nseq, st);
    if (e == cudaSuccess) {
        if (frame) e = launch_em_obs<float, 1>(st, m, b, ratios, nullptr, nullptr, nullptr, frame, seq_flag, sms);
        else if (prec == TEHMM_F32) e = launch_em_obs<float, 0>(st, m, b, ratios, (float *)elog, (float *)blin, rowmax, nullptr, seq_flag, sms);
        else e = launch_em_obs<double, 0>(st, m, b, ratios, (double *)elog, (double *)blin, rowmax, nullptr, seq_flag, sms);
    }
    if (e == cudaSuccess) {
        int warps = 4;
        int grid = (int)((b.nseq + warps - 1) / warps);
        if (frame || prec == TEHMM_F32)
            emission_fix_kernel<float><<<grid, warps * 32, 0, st>>>(m.N, m.LD, seq_flag, b.seq_off, b.nseq, (float *)elog, (float *)blin, rowmax, frame);
        else
            emission_fix_kernel<double><<<grid, warps * 32, 0, st>>>(m.N, m.LD, seq_flag, b.seq_off, b.nseq, (double *)elog, (double *)blin, rowmax, frame);
        e = cudaGetLastError();
    }
    *err = e;
    return e == cudaSuccess ? 2 : -1;
}
