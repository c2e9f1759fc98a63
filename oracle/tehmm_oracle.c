/*
 * tehmm_oracle.c -- CPU restatement of the teHmm multitrack-HMM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * product in tehmm_b200/.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * never links, imports or executes anything under oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here against (a) the known answers held by the reference's own tests
 * (tests/hmmTest.py, tests/emissionTest.py, tests/dpBenchmark.py) and
 * (b) outputs of the reference's own Cython modules, generated in the build
 * container by tests/golden/make_golden.py and committed under tests/golden/.
 *
 * Every function restates the arithmetic of one reference routine in the same
 * floating-point operation order (double precision, libm exp/log), so that the
 * results are expected to be bit-identical to the reference on the same libm.
 * Compile with -ffp-contract=off (see oracle/Makefile) to keep that property.
 *
 * Conventions: all matrices are dense row-major doubles.  `ratios` may be
 * NULL (the reference's segRatios=None).  obs is (T,K) row-major of
 * obs_bytes-wide unsigned/signed integers (1 = uint8, 2 = uint16, 4 = int32;
 * the three clones of _emission.pyx).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_ZEROLOGPROB (-1e200) /* _hmm.pyx:60 */
#define ORC_MINDBL (-1e20)       /* _emission.pyx:12 */

static inline long orc_sym(const void *obs, int obs_bytes, long idx)
{
    switch (obs_bytes) {
    case 1: return ((const uint8_t *)obs)[idx];
    case 2: return ((const uint16_t *)obs)[idx];
    default: return ((const int32_t *)obs)[idx];
    }
}

/* log(sum(exp(v))) with the max pulled out -- basehmm.py:70-93 (logsumexp). */
double orc_logsumexp(const double *v, long n)
{
    double vmax = v[0], acc = 0.0;
    for (long i = 1; i < n; ++i)
        if (v[i] > vmax) vmax = v[i];
    for (long i = 0; i < n; ++i)
        acc += exp(v[i] - vmax);
    return log(acc) + vmax;
}

/*
 * Emission gather-and-sum -- _emission.pyx:50-80 (and its U16 / int32 clones
 * at 82-144).  table is (K,N,S): [track][state][symbol].
 * The "impossible row" test uses a maximum that is carried across rows and
 * never reset (lines 59, 73-80): a row is zeroed only while no earlier entry
 * has exceeded -1e20.
 */
void orc_all_log_probs(const void *obs, int obs_bytes, long T, int K,
                       const double *table, int N, int S,
                       double *out, double normalize, const double *ratios)
{
    double running_max = ORC_MINDBL;
    for (long t = 0; t < T; ++t) {
        double *row = out + t * (long)N;
        for (int j = 0; j < N; ++j) {
            row[j] = 0.0;
            for (int k = 0; k < K; ++k)
                row[j] += table[((long)k * N + j) * S + orc_sym(obs, obs_bytes, t * K + k)];
            row[j] *= normalize;
            if (ratios) row[j] *= ratios[t];
            if (row[j] > running_max) running_max = row[j];
        }
        if (running_max == ORC_MINDBL)
            for (int j = 0; j < N; ++j) row[j] = 0.0;
    }
}

/*
 * Forward lattice -- _hmm.pyx:120-158.  The optional segment correction adds
 * logA[j][j]*(r_t-1) to every incoming term when r_t > 1 (lines 140-141,
 * 148-149).  Entries <= -1e200 become -inf (156-157).
 */
void orc_forward(long T, int N, const double *log_start, const double *log_trans,
                 const double *frame, const double *ratios, double *fwd)
{
    double *work = (double *)malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < N; ++i) {
        fwd[i] = log_start[i] + frame[i];
        if (ratios && ratios[0] > 1.)
            fwd[i] += log_trans[(long)i * N + i] * (ratios[0] - 1.);
    }
    for (long t = 1; t < T; ++t) {
        const double *prev = fwd + (t - 1) * N;
        double *cur = fwd + t * N;
        for (int j = 0; j < N; ++j) {
            double vmax = -INFINITY;
            for (int i = 0; i < N; ++i) {
                work[i] = prev[i] + log_trans[(long)i * N + j];
                if (ratios && ratios[t] > 1.)
                    work[i] += log_trans[(long)j * N + j] * (ratios[t] - 1.);
                if (work[i] > vmax) vmax = work[i];
            }
            double psum = 0.0;
            for (int i = 0; i < N; ++i) psum += exp(work[i] - vmax);
            cur[j] = log(psum) + vmax + frame[t * N + j];
            if (cur[j] <= ORC_ZEROLOGPROB) cur[j] = -INFINITY;
        }
    }
    free(work);
}

/*
 * Backward lattice -- _hmm.pyx:160-198.  The last row is log(1/N), not 0
 * (line 179).  log_start is accepted and ignored, as in the reference.
 */
void orc_backward(long T, int N, const double *log_start, const double *log_trans,
                  const double *frame, const double *ratios, double *bwd)
{
    (void)log_start;
    double *work = (double *)malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < N; ++i)
        bwd[(T - 1) * N + i] = log(1. / (double)N);
    for (long t = T - 2; t >= 0; --t) {
        const double *nxt = bwd + (t + 1) * N;
        const double *fr = frame + (t + 1) * N;
        for (int i = 0; i < N; ++i) {
            double vmax = -INFINITY;
            for (int j = 0; j < N; ++j) {
                work[j] = log_trans[(long)i * N + j] + fr[j] + nxt[j];
                if (ratios && ratios[t + 1] > 1.)
                    work[j] += log_trans[(long)j * N + j] * (ratios[t + 1] - 1.);
                if (work[j] > vmax) vmax = work[j];
            }
            double psum = 0.0;
            for (int j = 0; j < N; ++j) psum += exp(work[j] - vmax);
            bwd[t * N + i] = log(psum) + vmax;
            if (bwd[t * N + i] <= ORC_ZEROLOGPROB) bwd[t * N + i] = -INFINITY;
        }
    }
    free(work);
}

/*
 * Viterbi with back-pointers -- _hmm.pyx:201-259.
 * Tie rule: strict '>' scanning fromState upward, so the lowest fromState
 * wins (232-247); the end state is the first maximum (np.argmax, 252).
 * Segment quirk kept verbatim: for fromState 0 the term logA[j][j]*r_t is
 * added whenever ratios are present (no r_t>1 test) and logA[0][0] is taken
 * back out when j==0 (234-237); other fromStates add logA[j][j]*(r_t-1) only
 * when r_t>1 (243-244).
 * Returns the path log-probability; states receives T int64 values.
 */
double orc_viterbi(long T, int N, const double *log_start, const double *log_trans,
                   const double *ratios, const double *frame, int64_t *states)
{
    double *lat = (double *)calloc((size_t)T * N, sizeof(double));
    int16_t *bp = (int16_t *)malloc(sizeof(int16_t) * (size_t)T * N);
    for (int j = 0; j < N; ++j) {
        lat[j] = log_start[j] + frame[j];
        if (ratios && ratios[0] > 1.)
            lat[j] += log_trans[(long)j * N + j] * (ratios[0] - 1.);
    }
    for (long t = 1; t < T; ++t) {
        const double *prev = lat + (t - 1) * N;
        for (int j = 0; j < N; ++j) {
            double best = prev[0] + log_trans[j] + frame[t * N + j];
            if (ratios) {
                best += log_trans[(long)j * N + j] * ratios[t];
                if (j == 0) best -= log_trans[j];
            }
            int16_t arg = 0;
            for (int i = 1; i < N; ++i) {
                double cand = prev[i] + log_trans[(long)i * N + j] + frame[t * N + j];
                if (ratios && ratios[t] > 1.)
                    cand += log_trans[(long)j * N + j] * (ratios[t] - 1.);
                if (cand > best) { best = cand; arg = (int16_t)i; }
            }
            lat[t * N + j] = best;
            bp[t * N + j] = arg;
        }
    }
    long last = 0;
    for (int j = 1; j < N; ++j)
        if (lat[(T - 1) * N + j] > lat[(T - 1) * N + last]) last = j;
    double logprob = lat[(T - 1) * N + last];
    states[T - 1] = last;
    for (long t = T - 1; t > 0; --t)
        states[t - 1] = bp[t * N + states[t]];
    free(lat);
    free(bp);
    return logprob;
}

/*
 * Two-pass log-sum of the transition posteriors over time --
 * _hmm.pyx:62-117.  out must be zero on entry (it is accumulated with +=
 * before the final log, lines 111-117).  With ratios, transitions into a
 * segment longer than the effective length also deposit the (r-1) implied
 * self transitions on the diagonal (89-96, 106-111).
 */
void orc_log_sum_lneta(long T, int N, const double *fwd, const double *log_trans,
                       const double *bwd, const double *frame, double logprob,
                       const double *ratios, double *out)
{
    double *mx = (double *)malloc(sizeof(double) * (size_t)N * N);
    for (long e = 0; e < (long)N * N; ++e) mx[e] = -INFINITY;
    for (int pass = 0; pass < 2; ++pass) {
        for (long t = 0; t + 1 < T; ++t) {
            for (int i = 0; i < N; ++i) {
                for (int j = 0; j < N; ++j) {
                    long e = (long)i * N + j;
                    double x = fwd[t * N + i] + log_trans[e] + frame[(t + 1) * N + j]
                             + bwd[(t + 1) * N + j] - logprob;
                    if (ratios && ratios[t + 1] > 1.) {
                        x += log_trans[(long)j * N + j] * (ratios[t + 1] - 1.);
                        if (i == j) {
                            double y = fwd[(t + 1) * N + i] + bwd[(t + 1) * N + j]
                                     + log(ratios[t + 1] - 1.) - logprob;
                            if (pass == 0) { if (y > mx[e]) mx[e] = y; }
                            else out[e] += exp(y - mx[e]);
                        }
                    }
                    if (pass == 0) { if (x > mx[e]) mx[e] = x; }
                    else out[e] += exp(x - mx[e]);
                }
            }
        }
    }
    for (long e = 0; e < (long)N * N; ++e) out[e] = log(out[e]) + mx[e];
    free(mx);
}

/*
 * Posterior glue used by fit() -- basehmm.py:516-517:
 *   gamma = fwd + bwd ; post = exp(gamma - logsumexp_row(gamma)).
 * With renorm_eps != 0 also the score_samples tail (basehmm.py:271-272):
 *   post += float32 eps ; post /= rowsum.
 */
void orc_posteriors(long T, int N, const double *fwd, const double *bwd,
                    double *post, int renorm_eps)
{
    const double eps32 = 1.1920928955078125e-07;
    for (long t = 0; t < T; ++t) {
        double *row = post + t * N;
        for (int j = 0; j < N; ++j) row[j] = fwd[t * N + j] + bwd[t * N + j];
        double lse = orc_logsumexp(row, N);
        for (int j = 0; j < N; ++j) row[j] = exp(row[j] - lse);
        if (renorm_eps) {
            double s = 0.0;
            for (int j = 0; j < N; ++j) { row[j] += eps32; }
            /* numpy pairwise-sums rows of < 8 elements sequentially; for wider
             * rows the difference is at the 1e-16 level and the tests use a
             * tolerance there. */
            for (int j = 0; j < N; ++j) s += row[j];
            for (int j = 0; j < N; ++j) row[j] /= s;
        }
    }
}

/*
 * Posterior-weighted per-track histograms -- _emission.pyx:171-190 (and the
 * clones at 193-234).  stats is (K,N,S) and is accumulated in place.
 */
void orc_accumulate_stats(const void *obs, int obs_bytes, long T, int K,
                          double *stats, int N, int S,
                          const double *post, const double *ratios)
{
    for (long t = 0; t < T; ++t)
        for (int k = 0; k < K; ++k) {
            long sym = orc_sym(obs, obs_bytes, t * K + k);
            for (int j = 0; j < N; ++j) {
                double w = post[t * N + j];
                if (ratios) w *= ratios[t];
                stats[((long)k * N + j) * S + sym] += w;
            }
        }
}

/*
 * Supervised counts over [start,end) in table coordinates --
 * _emission.pyx:266-332.  Each position adds 1 (or ratios[pos]) to
 * stats[track][state][obs[pos][track]].
 */
void orc_update_counts(const void *obs, int obs_bytes, int K, long start, long end,
                       int state, double *stats, int N, int S, const double *ratios)
{
    for (long pos = start; pos < end; ++pos) {
        double w = ratios ? ratios[pos] : 1.0;
        for (int k = 0; k < K; ++k)
            stats[((long)k * N + state) * S + orc_sym(obs, obs_bytes, pos * K + k)] += w;
    }
}

/*
 * One sequence of the E-step exactly as BaseHMM.fit drives it
 * (basehmm.py:509-522 with hmm.py:545-574): emission, forward, backward,
 * posterior glue, start += post[0], trans += exp(logsum_lneta),
 * emission histograms.  Used as the timed "port" CPU baseline and as the
 * checker for the fused device E-step.  Returns the forward log-likelihood.
 * All stats are accumulated in place (obs_stats is (K,N,stats_S): initStats,
 * emission.py:212-214, is one wider than the table when zeroAsMissingData is off).
 * Scratch is allocated internally.
 */
double orc_estep_sequence(const void *obs, int obs_bytes, long T, int K,
                          const double *table, int N, int S, double normalize,
                          const double *log_start, const double *log_trans,
                          const double *ratios,
                          double *start_stats, double *trans_stats, double *obs_stats,
                          int stats_S)
{
    size_t cells = (size_t)T * N;
    double *frame = (double *)malloc(sizeof(double) * cells);
    double *fwd = (double *)calloc(cells, sizeof(double));
    double *bwd = (double *)calloc(cells, sizeof(double));
    double *post = (double *)malloc(sizeof(double) * cells);
    orc_all_log_probs(obs, obs_bytes, T, K, table, N, S, frame, normalize, ratios);
    orc_forward(T, N, log_start, log_trans, frame, ratios, fwd);
    double lp = orc_logsumexp(fwd + (T - 1) * N, N);
    orc_backward(T, N, log_start, log_trans, frame, ratios, bwd);
    orc_posteriors(T, N, fwd, bwd, post, 0);
    for (int j = 0; j < N; ++j) start_stats[j] += post[j];
    if (T > 1) {
        double *ls = (double *)calloc((size_t)N * N, sizeof(double));
        orc_log_sum_lneta(T, N, fwd, log_trans, bwd, frame, lp, ratios, ls);
        for (long e = 0; e < (long)N * N; ++e) trans_stats[e] += exp(ls[e]);
        free(ls);
    }
    orc_accumulate_stats(obs, obs_bytes, T, K, obs_stats, N, stats_S, post, ratios);
    free(frame); free(fwd); free(bwd); free(post);
    return lp;
}

/*
 * Decode sweep of one sequence: emission + forward + backward + posterior
 * argmax (MAP, basehmm.py:356-358) + Viterbi.  The timed "port" CPU baseline
 * for the fwd-bwd+Viterbi metric.  map_states / vit_states receive T int64.
 */
void orc_sweep_sequence(const void *obs, int obs_bytes, long T, int K,
                        const double *table, int N, int S, double normalize,
                        const double *log_start, const double *log_trans,
                        const double *ratios,
                        int64_t *vit_states, int64_t *map_states, double *out3)
{
    size_t cells = (size_t)T * N;
    double *frame = (double *)malloc(sizeof(double) * cells);
    double *fwd = (double *)calloc(cells, sizeof(double));
    double *bwd = (double *)calloc(cells, sizeof(double));
    orc_all_log_probs(obs, obs_bytes, T, K, table, N, S, frame, normalize, ratios);
    orc_forward(T, N, log_start, log_trans, frame, ratios, fwd);
    out3[0] = orc_logsumexp(fwd + (T - 1) * N, N);
    orc_backward(T, N, log_start, log_trans, frame, ratios, bwd);
    orc_posteriors(T, N, fwd, bwd, fwd, 1);
    double acc = 0.0;
    for (long t = 0; t < T; ++t) {
        int best = 0;
        for (int j = 1; j < N; ++j)
            if (fwd[t * N + j] > fwd[t * N + best]) best = j;
        map_states[t] = best;
        acc += fwd[t * N + best];
    }
    out3[1] = acc;
    out3[2] = orc_viterbi(T, N, log_start, log_trans, ratios, frame, vit_states);
    free(frame); free(fwd); free(bwd);
}
