// L0 "strict" kernels: float64, one sequence per call, the reference's
// operation order (compile with -fmad=false so no multiply-add is contracted).
//
// Each kernel cites the reference routine it replaces.  Adds, multiplies and
// comparisons are IEEE-exact and in the reference's order, so Viterbi paths,
// emission sums and count histograms are bit-identical to the reference;
// forward/backward/lneta differ only by CUDA libdevice exp/log vs glibc
// (<= 1 ulp per call).
#include "common.cuh"

// ---------------------------------------------------------------- emission
// _emission.pyx:50-80 (+ clones 82-144): out[t][j] = ((sum_k table[k][j][obs[t][k]])
// * normalize) * ratio[t].  The running-max quirk (lines 59, 73-80) is applied
// by strict_emission_fix_kernel once the first feasible row is known.
template <typename OBS>
__global__ void strict_emission_kernel(const OBS *__restrict__ obs, int64_t T, int K,
                                       const double *__restrict__ table, int N, int S,
                                       double *__restrict__ out, double normalize,
                                       const double *__restrict__ ratios,
                                       unsigned long long *first_feasible)
{
    int64_t cells = T * (int64_t)N;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < cells;
         c += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = c / N;
        int j = (int)(c - t * N);
        double v = 0.0;
        for (int k = 0; k < K; ++k)
            v += table[((int64_t)k * N + j) * S + (int64_t)obs[t * K + k]];
        v *= normalize;
        if (ratios) v *= ratios[t];
        out[c] = v;
        if (v > TEHMM_MINDBL) atomicMin(first_feasible, (unsigned long long)t);
    }
}

__global__ void strict_emission_fix_kernel(double *out, int64_t T, int N,
                                           const unsigned long long *first_feasible)
{
    int64_t lim = (int64_t)min((unsigned long long)T, *first_feasible) * N;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < lim;
         c += (int64_t)gridDim.x * blockDim.x)
        out[c] = 0.0;
}

// ---------------------------------------------------------------- forward
// _hmm.pyx:120-158.  One CTA; thread j owns states j, j+blockDim, ...
__global__ void strict_forward_kernel(int64_t T, int N, const double *__restrict__ log_start,
                                      const double *__restrict__ log_trans,
                                      const double *__restrict__ frame,
                                      const double *__restrict__ ratios,
                                      double *__restrict__ fwd)
{
    extern __shared__ double sh[];
    double *prev = sh, *cur = sh + N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double v = log_start[j] + frame[j];
        if (ratios && ratios[0] > 1.) v += log_trans[(int64_t)j * N + j] * (ratios[0] - 1.);
        prev[j] = v;
        fwd[j] = v;
    }
    __syncthreads();
    for (int64_t t = 1; t < T; ++t) {
        double r = ratios ? ratios[t] : 0.0;
        bool seg = ratios && r > 1.;
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            double ajj = log_trans[(int64_t)j * N + j];
            double vmax = -INFINITY;
            for (int i = 0; i < N; ++i) {
                double w = prev[i] + log_trans[(int64_t)i * N + j];
                if (seg) w += ajj * (r - 1.);
                if (w > vmax) vmax = w;
            }
            double psum = 0.0;
            for (int i = 0; i < N; ++i) {
                double w = prev[i] + log_trans[(int64_t)i * N + j];
                if (seg) w += ajj * (r - 1.);
                psum += exp(w - vmax);
            }
            double v = log(psum) + vmax + frame[t * N + j];
            if (v <= TEHMM_ZEROLOGPROB) v = -INFINITY;
            cur[j] = v;
            fwd[t * N + j] = v;
        }
        __syncthreads();
        double *tmp = prev; prev = cur; cur = tmp;
    }
}

// ---------------------------------------------------------------- backward
// _hmm.pyx:160-198.  Last row is log(1/N) (line 179).
__global__ void strict_backward_kernel(int64_t T, int N, const double *__restrict__ log_trans,
                                       const double *__restrict__ frame,
                                       const double *__restrict__ ratios,
                                       double *__restrict__ bwd)
{
    extern __shared__ double sh[];
    double *nxt = sh, *cur = sh + N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double v = log(1. / (double)N);
        nxt[i] = v;
        bwd[(T - 1) * N + i] = v;
    }
    __syncthreads();
    for (int64_t t = T - 2; t >= 0; --t) {
        double r = ratios ? ratios[t + 1] : 0.0;
        bool seg = ratios && r > 1.;
        const double *fr = frame + (t + 1) * N;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            double vmax = -INFINITY;
            for (int j = 0; j < N; ++j) {
                double w = log_trans[(int64_t)i * N + j] + fr[j] + nxt[j];
                if (seg) w += log_trans[(int64_t)j * N + j] * (r - 1.);
                if (w > vmax) vmax = w;
            }
            double psum = 0.0;
            for (int j = 0; j < N; ++j) {
                double w = log_trans[(int64_t)i * N + j] + fr[j] + nxt[j];
                if (seg) w += log_trans[(int64_t)j * N + j] * (r - 1.);
                psum += exp(w - vmax);
            }
            double v = log(psum) + vmax;
            if (v <= TEHMM_ZEROLOGPROB) v = -INFINITY;
            cur[i] = v;
            bwd[t * N + i] = v;
        }
        __syncthreads();
        double *tmp = nxt; nxt = cur; cur = tmp;
    }
}

// ---------------------------------------------------------------- viterbi
// _hmm.pyx:201-259 incl. the fromState==0 segment quirk (234-237) and the
// strict '>' / first-maximum tie rules (245, 252).
__global__ void strict_viterbi_kernel(int64_t T, int N, const double *__restrict__ log_start,
                                      const double *__restrict__ log_trans,
                                      const double *__restrict__ ratios,
                                      const double *__restrict__ frame,
                                      int16_t *__restrict__ bp, int64_t *__restrict__ states,
                                      double *__restrict__ logprob)
{
    extern __shared__ double sh[];
    double *prev = sh, *cur = sh + N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double v = log_start[j] + frame[j];
        if (ratios && ratios[0] > 1.) v += log_trans[(int64_t)j * N + j] * (ratios[0] - 1.);
        prev[j] = v;
    }
    __syncthreads();
    for (int64_t t = 1; t < T; ++t) {
        double r = ratios ? ratios[t] : 0.0;
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            double ajj = log_trans[(int64_t)j * N + j];
            double b = frame[t * N + j];
            double best = prev[0] + log_trans[j] + b;
            if (ratios) {
                best += ajj * r;
                if (j == 0) best -= log_trans[j];
            }
            int arg = 0;
            for (int i = 1; i < N; ++i) {
                double cand = prev[i] + log_trans[(int64_t)i * N + j] + b;
                if (ratios && r > 1.) cand += ajj * (r - 1.);
                if (cand > best) { best = cand; arg = i; }
            }
            cur[j] = best;
            bp[t * N + j] = (int16_t)arg;
        }
        __syncthreads();
        double *tmp = prev; prev = cur; cur = tmp;
    }
    if (threadIdx.x == 0) {
        int last = 0;
        for (int j = 1; j < N; ++j)
            if (prev[j] > prev[last]) last = j;
        *logprob = prev[last];
        int64_t s = last;
        states[T - 1] = s;
        for (int64_t t = T - 1; t > 0; --t) {
            s = bp[t * N + s];
            states[t - 1] = s;
        }
    }
}

// ---------------------------------------------------------------- lneta
// _hmm.pyx:62-117.  One thread per (i,j); two passes over t in order.
__global__ void strict_lneta_kernel(int64_t T, int N, const double *__restrict__ fwd,
                                    const double *__restrict__ log_trans,
                                    const double *__restrict__ bwd,
                                    const double *__restrict__ frame, double logprob,
                                    const double *__restrict__ ratios, double *__restrict__ out)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * N) return;
    int i = e / N, j = e - i * N;
    double aij = log_trans[e];
    double ajj = log_trans[(int64_t)j * N + j];
    double mx = -INFINITY;
    for (int64_t t = 0; t + 1 < T; ++t) {
        double x = fwd[t * N + i] + aij + frame[(t + 1) * N + j] + bwd[(t + 1) * N + j] - logprob;
        if (ratios && ratios[t + 1] > 1.) {
            x += ajj * (ratios[t + 1] - 1.);
            if (i == j) {
                double y = fwd[(t + 1) * N + i] + bwd[(t + 1) * N + j] + log(ratios[t + 1] - 1.) - logprob;
                if (y > mx) mx = y;
            }
        }
        if (x > mx) mx = x;
    }
    double acc = out[e];
    for (int64_t t = 0; t + 1 < T; ++t) {
        double x = fwd[t * N + i] + aij + frame[(t + 1) * N + j] + bwd[(t + 1) * N + j] - logprob;
        if (ratios && ratios[t + 1] > 1.) {
            x += ajj * (ratios[t + 1] - 1.);
            if (i == j) {
                double y = fwd[(t + 1) * N + i] + bwd[(t + 1) * N + j] + log(ratios[t + 1] - 1.) - logprob;
                acc += exp(y - mx);
            }
        }
        acc += exp(x - mx);
    }
    out[e] = log(acc) + mx;
}

// ---------------------------------------------------------------- statistics
// _emission.pyx:171-190 (+ clones).  One thread per (track,state), t in order.
template <typename OBS>
__global__ void strict_accumulate_kernel(const OBS *__restrict__ obs, int64_t T, int K,
                                         double *__restrict__ stats, int N, int S,
                                         const double *__restrict__ post,
                                         const double *__restrict__ ratios)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K * N) return;
    int k = e / N, j = e - k * N;
    double *row = stats + (int64_t)e * S;
    for (int64_t t = 0; t < T; ++t) {
        double w = post[t * N + j];
        if (ratios) w *= ratios[t];
        row[(int64_t)obs[t * K + k]] += w;
    }
}

// _emission.pyx:266-332.  One thread per track, positions in order.
template <typename OBS>
__global__ void strict_counts_kernel(const OBS *__restrict__ obs, int K, int64_t start,
                                     int64_t end, int state, double *__restrict__ stats,
                                     int N, int S, const double *__restrict__ ratios)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double *row = stats + ((int64_t)k * N + state) * S;
    for (int64_t pos = start; pos < end; ++pos) {
        double w = ratios ? ratios[pos] : 1.0;
        row[(int64_t)obs[pos * K + k]] += w;
    }
}

// ---------------------------------------------------------------- launchers
static inline int strict_threads(int N) { int t = ((N + 31) / 32) * 32; return t > 1024 ? 1024 : t; }

#define OBS_DISPATCH(BYTES, CALL)                                              \
    do {                                                                       \
        if ((BYTES) == 1) { typedef uint8_t OBS_T; CALL; }                     \
        else if ((BYTES) == 2) { typedef uint16_t OBS_T; CALL; }               \
        else { typedef int32_t OBS_T; CALL; }                                  \
    } while (0)

void tehmm_launch_strict_emission(cudaStream_t st, const void *obs, int obs_bytes, int64_t T,
                                  int K, const double *table, int N, int S, double *out,
                                  double normalize, const double *ratios,
                                  unsigned long long *d_first)
{
    int64_t cells = T * (int64_t)N;
    int blocks = (int)((cells + 255) / 256 > 148 * 16 ? 148 * 16 : (cells + 255) / 256);
    if (blocks < 1) blocks = 1;
    OBS_DISPATCH(obs_bytes, (strict_emission_kernel<OBS_T><<<blocks, 256, 0, st>>>(
        (const OBS_T *)obs, T, K, table, N, S, out, normalize, ratios, d_first)));
    strict_emission_fix_kernel<<<blocks, 256, 0, st>>>(out, T, N, d_first);
}

void tehmm_launch_strict_forward(cudaStream_t st, int64_t T, int N, const double *ls,
                                 const double *lt, const double *frame, const double *ratios,
                                 double *fwd)
{
    strict_forward_kernel<<<1, strict_threads(N), 2 * N * sizeof(double), st>>>(T, N, ls, lt, frame, ratios, fwd);
}

void tehmm_launch_strict_backward(cudaStream_t st, int64_t T, int N, const double *lt,
                                  const double *frame, const double *ratios, double *bwd)
{
    strict_backward_kernel<<<1, strict_threads(N), 2 * N * sizeof(double), st>>>(T, N, lt, frame, ratios, bwd);
}

void tehmm_launch_strict_viterbi(cudaStream_t st, int64_t T, int N, const double *ls,
                                 const double *lt, const double *ratios, const double *frame,
                                 int16_t *bp, int64_t *states, double *logprob)
{
    strict_viterbi_kernel<<<1, strict_threads(N), 2 * N * sizeof(double), st>>>(T, N, ls, lt, ratios, frame, bp, states, logprob);
}

void tehmm_launch_strict_lneta(cudaStream_t st, int64_t T, int N, const double *fwd,
                               const double *lt, const double *bwd, const double *frame,
                               double logprob, const double *ratios, double *out)
{
    int cells = N * N;
    strict_lneta_kernel<<<(cells + 63) / 64, 64, 0, st>>>(T, N, fwd, lt, bwd, frame, logprob, ratios, out);
}

void tehmm_launch_strict_accumulate(cudaStream_t st, const void *obs, int obs_bytes, int64_t T,
                                    int K, double *stats, int N, int S, const double *post,
                                    const double *ratios)
{
    int cells = K * N;
    OBS_DISPATCH(obs_bytes, (strict_accumulate_kernel<OBS_T><<<(cells + 31) / 32, 32, 0, st>>>(
        (const OBS_T *)obs, T, K, stats, N, S, post, ratios)));
}

void tehmm_launch_strict_counts(cudaStream_t st, const void *obs, int obs_bytes, int K,
                                int64_t start, int64_t end, int state, double *stats, int N,
                                int S, const double *ratios)
{
    OBS_DISPATCH(obs_bytes, (strict_counts_kernel<OBS_T><<<(K + 31) / 32, 32, 0, st>>>(
        (const OBS_T *)obs, K, start, end, state, stats, N, S, ratios)));
}
