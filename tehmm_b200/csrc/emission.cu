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

// ---------------------------------------------------------------------------
// fp32 production path: merged, centred tables (TehmmModelDev::gtab / gc).
// The generic kernel above is bound by shared-memory bandwidth: K float64 table
// rows of 8N bytes per time step (ncu: 1.07 ms of LDS wavefronts at 10 M x 30 x 10).
// Here tracks are merged into G <= K groups (one look-up per group) and a table
// row is 128 bytes of float32 -- the part of the log-probability that differs
// between states -- while the part common to all states (the row's maximum, kept
// in float64) is summed per time step by one lane and goes into rowmax.  The
// float32 rounding is that of the elog lattice itself (relative 6e-8 of the
// distance to the row maximum); the float64 verification mode never takes this
// path.  One warp per time step, lane = state, 16 steps staged per iteration.
#define EMG_WARPS 32
#define EMG_ROWS 16

__device__ __forceinline__ void emg_flag_row(int64_t t, const int64_t *seq_off, int64_t nseq, int *seq_flag)
{
    int64_t lo = 0, hi = nseq;   // sequence of row t
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (seq_off[mid] <= t) lo = mid; else hi = mid; }
    seq_flag[lo] = 1;
}

// GT = number of groups when 1..8 (look-ups fully unrolled), 0 = any (loop).
// NS = 1: N <= 32 (LD = 32); NS = 2: 33..64 states (LD = N, table rows of 64 floats,
// lane owns columns lane and lane + 32).
template <typename OBS, int GT, bool RATIO, int NS>
__global__ void __launch_bounds__(EMG_WARPS * 32, 1)
emission_merged_kernel(TehmmModelDev m, const OBS *__restrict__ obs, int64_t total,
                       const double *__restrict__ ratios, float *__restrict__ elog,
                       float *__restrict__ blin, double *__restrict__ rowmax,
                       int *__restrict__ seq_flag, const int64_t *__restrict__ seq_off, int64_t nseq)
{
    extern __shared__ __align__(16) unsigned char em_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = m.K, N = m.N, LD = m.LD;
    const int G = GT ? GT : m.G;
    constexpr int NP = 32 * NS;
    constexpr int ROWB = NP * 4;            // bytes of a merged table row
    constexpr int ROWSH = NS == 1 ? 7 : 8;  // log2(ROWB)
    // layout: gtab | gc | gdesc | nsym | per warp { csum[16], offs[16][16] }
    float *tab_s = reinterpret_cast<float *>(em_smem);
    double *gc_s = reinterpret_cast<double *>(tab_s + (size_t)m.grows * NP);
    int32_t *gd_s = reinterpret_cast<int32_t *>(gc_s + m.grows);
    int32_t *nsym_s = gd_s + TEHMM_GMAX * TEHMM_GDESC;
    const size_t per_warp = (size_t)EMG_ROWS * 8 + (size_t)EMG_ROWS * 16 * 4;
    const size_t head = ((size_t)m.grows * (ROWB + 8) + (size_t)(TEHMM_GMAX * TEHMM_GDESC + K) * 4 + 15) & ~(size_t)15;
    unsigned char *wbase = em_smem + head + (size_t)warp * per_warp;
    int32_t *offs = reinterpret_cast<int32_t *>(wbase + EMG_ROWS * 8);

    for (int64_t e = threadIdx.x; e < (int64_t)m.grows * NP; e += blockDim.x) tab_s[e] = m.gtab[e];
    for (int e = threadIdx.x; e < m.grows; e += blockDim.x) gc_s[e] = m.gc[e];
    for (int e = threadIdx.x; e < G * TEHMM_GDESC; e += blockDim.x) gd_s[e] = m.gdesc[e];
    for (int e = threadIdx.x; e < K; e += blockDim.x) nsym_s[e] = m.track_nsym[e];
    __syncthreads();

    // shared-space address of this lane's column of table row 0; offs holds byte offsets of rows
    const uint32_t lane_tab = (uint32_t)__cvta_generic_to_shared(tab_s) + (uint32_t)lane * 4u;
    const uint32_t offs_a = (uint32_t)__cvta_generic_to_shared(offs);
    bool is_state[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) is_state[s] = lane + 32 * s < N;

    // state-dependent part of row r of the staged batch: sum of the groups' table rows (all <= 0)
    auto gather = [&](int r, float (&v)[NS]) {
#pragma unroll
        for (int s = 0; s < NS; ++s) v[s] = 0.f;
        if (GT) {
            int o[8];
            asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(offs_a + r * 64));
            if (GT > 4)
                asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]) : "r"(offs_a + r * 64 + 16));
#pragma unroll
            for (int gq = 0; gq < GT; ++gq) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    float x;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(lane_tab + (uint32_t)o[gq] + 128u * s));
                    v[s] += x;
                }
            }
        } else {
            for (int gq = 0; gq < G; ++gq) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    float x;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(lane_tab + (uint32_t)offs[r * 16 + gq] + 128u * s));
                    v[s] += x;
                }
            }
        }
    };
    // normalised row out (pe / pb: this lane's first column of the row); returns the maximum
    // over states taken out of v
    auto emit_row = [&](float *pe, float *pb, const float (&v)[NS], float rf) -> float {
        // v <= 0 everywhere: the bit patterns of non-positive floats grow with the magnitude,
        // so the maximum is the unsigned minimum (+0.0 = 0 included; -inf is the largest)
        unsigned bits = 0xff800000u;
#pragma unroll
        for (int s = 0; s < NS; ++s) bits = min(bits, is_state[s] ? __float_as_uint(v[s]) : 0xff800000u);
        const float Mf = __uint_as_float(__reduce_min_sync(TEHMM_FULL, bits));
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            float d = Mf > -INFINITY ? v[s] - Mf : 0.f;
            if (RATIO) d *= rf;
            float bl = __expf(d);
            if (!is_state[s]) { d = 0.f; bl = 0.f; }          // padding columns
            if (lane + 32 * s < LD) {
                if (pe) pe[32 * s] = d;
                if (pb) pb[32 * s] = bl;
            }
        }
        return Mf;
    };

    const int64_t nblocks = (total + EMG_ROWS - 1) / EMG_ROWS;
    for (int64_t blk = (int64_t)blockIdx.x * EMG_WARPS + warp; blk < nblocks;
         blk += (int64_t)gridDim.x * EMG_WARPS) {
        const int64_t tb = blk * EMG_ROWS;
        const int rows = (int)min((int64_t)EMG_ROWS, total - tb);
        __syncwarp();
        // ---- stage: byte offset of the table row of every (row, group); symbols straight from HBM
        bool bad = false;
        for (int e = lane; e < rows * G; e += 32) {
            const int r = e / G, gq = e - r * G;
            const int32_t *d = gd_s + gq * TEHMM_GDESC;
            int idx = d[1];
            for (int i = 0; i < d[0]; ++i) {
                const int k = d[2 + i];
                const int sym = (int)obs[(tb + r) * K + k];
                bad |= (unsigned)sym >= (unsigned)nsym_s[k];
                idx += sym * d[6 + i];
            }
            offs[r * 16 + gq] = idx * ROWB;
        }
        const bool any_slow = __any_sync(TEHMM_FULL, bad);
        __syncwarp();
        float mf_keep = 0.f;                     // lane r keeps the state maximum of row r
        double c_keep = 0.0;
        if (!any_slow) {
            if (lane < rows) {
                double c = 0.0;
                for (int gq = 0; gq < G; ++gq) c += gc_s[offs[lane * 16 + gq] >> ROWSH];
                c_keep = c;
            }
            float *pe = elog ? elog + tb * LD + lane : nullptr;
            float *pb = blin ? blin + tb * LD + lane : nullptr;
            if (rows == EMG_ROWS) {
#pragma unroll
                for (int r = 0; r < EMG_ROWS; r += 2) {          // two rows in flight
                    float v0[NS], v1[NS];
                    gather(r, v0);
                    gather(r + 1, v1);
                    const float r0 = RATIO ? (float)ratios[tb + r] : 1.f, r1 = RATIO ? (float)ratios[tb + r + 1] : 1.f;
                    const float M0 = emit_row(pe ? pe + r * LD : nullptr, pb ? pb + r * LD : nullptr, v0, r0);
                    const float M1 = emit_row(pe ? pe + (r + 1) * LD : nullptr, pb ? pb + (r + 1) * LD : nullptr, v1, r1);
                    if (lane == r) mf_keep = M0;
                    if (lane == r + 1) mf_keep = M1;
                }
            } else {
                for (int r = 0; r < rows; ++r) {
                    float v[NS];
                    gather(r, v);
                    const float Mr = emit_row(pe ? pe + r * LD : nullptr, pb ? pb + r * LD : nullptr, v,
                                              RATIO ? (float)ratios[tb + r] : 1.f);
                    if (lane == r) mf_keep = Mr;
                }
            }
        } else {
            // a symbol outside its track's table: index the dense float64 table exactly as the
            // reference does (rare)
            for (int r = 0; r < rows; ++r) {
                double v[NS], vm = -INFINITY;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    v[s] = 0.0;
                    if (is_state[s])
                        for (int k = 0; k < K; ++k) v[s] += m.table[((int64_t)k * N + lane + 32 * s) * m.S + (int64_t)obs[(tb + r) * K + k]];
                    v[s] *= m.normalize;
                    if (is_state[s]) vm = fmax(vm, v[s]);
                }
                const double M = warp_max(vm);
                float vf[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) vf[s] = M > -INFINITY ? (float)(v[s] - M) : 0.f;
                emit_row(elog ? elog + (tb + r) * LD + lane : nullptr, blin ? blin + (tb + r) * LD + lane : nullptr,
                         vf, RATIO ? (float)ratios[tb + r] : 1.f);
                if (lane == r) { mf_keep = 0.f; c_keep = M; }
            }
        }
        // ---- row maxima of the batch, one lane per row
        if (lane < rows) {
            double M = (double)mf_keep + c_keep;
            if (RATIO) M *= ratios[tb + lane];
            rowmax[tb + lane] = M;
            if (!(M > TEHMM_MINDBL) && seq_flag) emg_flag_row(tb + lane, seq_off, nseq, seq_flag);   // _emission.pyx:73-80
        }
    }
}

static size_t emg_smem_bytes(const TehmmModelDev &m)
{
    const size_t per_warp = (size_t)EMG_ROWS * 8 + (size_t)EMG_ROWS * 16 * 4;
    const size_t head = ((size_t)m.grows * (128 * m.NS + 8) + (size_t)(TEHMM_GMAX * TEHMM_GDESC + m.K) * 4 + 15) & ~(size_t)15;
    return head + EMG_WARPS * per_warp;
}

template <typename OBS, int GT, bool RATIO, int NS>
static cudaError_t launch_emg3(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                               const double *ratios, float *elog, float *blin, double *rowmax,
                               int *seq_flag, int sms)
{
    const size_t smem = emg_smem_bytes(m);
    cudaError_t e = cudaFuncSetAttribute(emission_merged_kernel<OBS, GT, RATIO, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t need = ((b.total + EMG_ROWS - 1) / EMG_ROWS + EMG_WARPS - 1) / EMG_WARPS;
    if (need < 1) need = 1;
    const int grid = (int)(need < sms ? need : sms);
    emission_merged_kernel<OBS, GT, RATIO, NS><<<grid, EMG_WARPS * 32, smem, st>>>(m, (const OBS *)b.obs, b.total, ratios, elog, blin,
                                                                                   rowmax, seq_flag, b.seq_off, b.nseq);
    return cudaGetLastError();
}

template <typename OBS>
static cudaError_t launch_emg(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                              const double *ratios, float *elog, float *blin, double *rowmax,
                              int *seq_flag, int sms)
{
#define EMG_GO1(GT, NS_) (ratios ? launch_emg3<OBS, GT, true, NS_>(st, m, b, ratios, elog, blin, rowmax, seq_flag, sms) \
                                 : launch_emg3<OBS, GT, false, NS_>(st, m, b, ratios, elog, blin, rowmax, seq_flag, sms))
#define EMG_GO(GT) (m.NS == 1 ? EMG_GO1(GT, 1) : EMG_GO1(GT, 2))
    switch (m.G) {
    case 1: return EMG_GO(1);
    case 2: return EMG_GO(2);
    case 3: return EMG_GO(3);
    case 4: return EMG_GO(4);
    case 5: return EMG_GO(5);
    case 6: return EMG_GO(6);
    case 7: return EMG_GO(7);
    case 8: return EMG_GO(8);
    default: return EMG_GO(0);
    }
#undef EMG_GO
#undef EMG_GO1
}

// ---------------------------------------------------------------------------
// Four time steps per warp pass (fp32, LD == 32, G <= 8): the production kernel.
// emission_merged_kernel above spends one warp instruction per (row, group)
// look-up and per row reduction: 78 instructions per row, issue bound at 0.85 ms
// for 10 M rows against 0.42 ms of HBM time.  Here a row belongs to EIGHT lanes,
// each owning four consecutive states: a look-up is one LDS.128 serving four
// rows (each quarter-warp reads one 128-byte table row: 4 wavefronts, no
// conflicts), the row maximum is a 4-way local max + three xor-shuffles, and
// the outputs leave as STG.128 (four rows = 512 contiguous bytes per store).
// ~20 instructions per row; the sum over groups keeps the group order, so the
// values are bit-identical to emission_merged_kernel's.
// Table padding (states N..31) is -inf, so padding never wins the maximum.
#define EM4_WARPS 32
#define EM4_ROWS 32

static size_t em4_smem_bytes(const TehmmModelDev &m)
{
    const size_t head = ((size_t)m.grows * (128 * m.NS + 8) + (size_t)(TEHMM_GMAX * TEHMM_GDESC + m.K) * 4 + 15) & ~(size_t)15;
    return head + (size_t)EM4_WARPS * EM4_ROWS * 8 * 4;
}

// NS = 2 (33..64 states, lattice rows of 64 floats; round 2): a row belongs to SIXTEEN lanes, two rows per pass.
template <typename OBS, int GT, bool RATIO, int NS>
__global__ void __launch_bounds__(EM4_WARPS * 32, 1)
emission_merged4_kernel(TehmmModelDev m, const OBS *__restrict__ obs, int64_t total,
                        const double *__restrict__ ratios, float *__restrict__ elog,
                        float *__restrict__ blin, double *__restrict__ rowmax,
                        int *__restrict__ seq_flag, const int64_t *__restrict__ seq_off, int64_t nseq,
                        int64_t row0)
{
    // rows [row0, total) of the batch (row0 a multiple of EM4_ROWS): a caller that streams the
    // observations in can run the rows that have arrived (tehmm_run_emission_rows)
    extern __shared__ __align__(16) unsigned char em_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = m.K, N = m.N;
    // layout: gtab | gc | gdesc | nsym | per warp offs[32][8]
    float *tab_s = reinterpret_cast<float *>(em_smem);
    constexpr int NP = 32 * NS, LPR = 8 * NS, RPP = 32 / LPR;      // columns, lanes per row, rows per pass
    constexpr int ROWB = 128 * NS;                                  // bytes of a merged table row
    double *gc_s = reinterpret_cast<double *>(tab_s + (size_t)m.grows * NP);
    int32_t *gd_s = reinterpret_cast<int32_t *>(gc_s + m.grows);
    int32_t *nsym_s = gd_s + TEHMM_GMAX * TEHMM_GDESC;
    const size_t head = ((size_t)m.grows * (ROWB + 8) + (size_t)(TEHMM_GMAX * TEHMM_GDESC + K) * 4 + 15) & ~(size_t)15;
    int32_t *offs = reinterpret_cast<int32_t *>(em_smem + head) + (size_t)warp * EM4_ROWS * 8;

    for (int64_t e = threadIdx.x; e < (int64_t)m.grows * NP; e += blockDim.x) tab_s[e] = m.gtab[e];
    for (int e = threadIdx.x; e < m.grows; e += blockDim.x) gc_s[e] = m.gc[e];
    for (int e = threadIdx.x; e < GT * TEHMM_GDESC; e += blockDim.x) gd_s[e] = m.gdesc[e];
    for (int e = threadIdx.x; e < K; e += blockDim.x) nsym_s[e] = m.track_nsym[e];
    __syncthreads();

    const int q = lane / LPR, c = lane % LPR;
    const uint32_t lane_tab = (uint32_t)__cvta_generic_to_shared(tab_s) + (uint32_t)c * 16u;
    const uint32_t offs_a = (uint32_t)__cvta_generic_to_shared(offs);
    // pm[i]: -inf for a real state (max(d, -inf) = d), 0 for a padding column (max(-inf, 0) = 0)
    float pm[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pm[i] = (4 * c + i < N) ? -INFINITY : 0.f;

    const int64_t nblocks = (total + EM4_ROWS - 1) / EM4_ROWS;
    for (int64_t blk = row0 / EM4_ROWS + (int64_t)blockIdx.x * EM4_WARPS + warp; blk < nblocks;
         blk += (int64_t)gridDim.x * EM4_WARPS) {
        const int64_t tb = blk * EM4_ROWS;
        const int rows = (int)min((int64_t)EM4_ROWS, total - tb);
        __syncwarp();
        // ---- stage (lane = row): byte offset of the table row of every group, and the
        //      float64 part common to all states
        bool bad = false;
        double csum = 0.0;
        {
            const OBS *orow = obs + (tb + (lane < rows ? lane : 0)) * K;
#pragma unroll
            for (int gq = 0; gq < GT; ++gq) {
                const int32_t *d = gd_s + gq * TEHMM_GDESC;
                int idx = d[1];
                const int nt = d[0];
                for (int i = 0; i < nt; ++i) {
                    const int k = d[2 + i];
                    const int sym = (int)orow[k];
                    bad |= (unsigned)sym >= (unsigned)nsym_s[k];
                    idx += sym * d[6 + i];
                }
                if (bad) idx = 0;
                offs[lane * 8 + gq] = idx * ROWB;
                csum += gc_s[idx];
            }
        }
        const bool any_slow = __any_sync(TEHMM_FULL, bad && lane < rows);
        __syncwarp();
        float mf_keep = 0.f;                     // lane r keeps the state maximum of row r
        if (!any_slow) {
#pragma unroll 2
            for (int p = 0; p < EM4_ROWS / RPP; ++p) {
                const int r = RPP * p + q;
                int o[8];
                asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(offs_a + r * 32));
                if (GT > 4)
                    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]) : "r"(offs_a + r * 32 + 16));
                float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int gq = 0; gq < GT; ++gq) {
                    float x0, x1, x2, x3;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(lane_tab + (uint32_t)o[gq]));
                    v[0] += x0; v[1] += x1; v[2] += x2; v[3] += x3;
                }
                float mx = fmax3(v[0], v[1], fmaxf(v[2], v[3]));
                mx = fmaxf(mx, __shfl_xor_sync(TEHMM_FULL, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(TEHMM_FULL, mx, 2));
                mx = fmaxf(mx, __shfl_xor_sync(TEHMM_FULL, mx, 4));
                if (NS == 2) mx = fmaxf(mx, __shfl_xor_sync(TEHMM_FULL, mx, 8));
                const float rf = RATIO ? (float)ratios[min(tb + r, total - 1)] : 1.f;
                float d[4], bl[4];
                if (__builtin_expect(__any_sync(TEHMM_FULL, !(mx > -INFINITY)), 0)) {
                    // a row no state can emit (in float): zeros / ones, as emission_merged_kernel
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        d[i] = mx > -INFINITY ? v[i] - mx : 0.f;
                        if (RATIO) d[i] *= rf;
                        bl[i] = __expf(d[i]);
                        if (pm[i] == 0.f) { d[i] = 0.f; bl[i] = 0.f; }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        d[i] = v[i] - mx;
                        if (RATIO) d[i] *= rf;
                        bl[i] = __expf(d[i]);            // padding: exp(-inf) = 0
                        d[i] = fmaxf(d[i], pm[i]);       // padding: 0
                    }
                }
                if (r < rows) {
                    const int64_t o4 = (tb + r) * NP + 4 * c;
                    if (elog) *reinterpret_cast<float4 *>(elog + o4) = make_float4(d[0], d[1], d[2], d[3]);
                    if (blin) *reinterpret_cast<float4 *>(blin + o4) = make_float4(bl[0], bl[1], bl[2], bl[3]);
                }
                const float mm = __shfl_sync(TEHMM_FULL, mx, (lane % RPP) * LPR);      // lane r keeps row r
                if (lane / RPP == p) mf_keep = mm;
            }
        } else {
            // a symbol outside its track's table: index the dense float64 table exactly as the
            // reference does (rare); lane = state
            for (int r = 0; r < rows; ++r) {
                double v[NS], vm = -INFINITY;
#pragma unroll
                for (int s2 = 0; s2 < NS; ++s2) {
                    v[s2] = 0.0;
                    if (lane + 32 * s2 < N)
                        for (int k = 0; k < K; ++k) v[s2] += m.table[((int64_t)k * N + lane + 32 * s2) * m.S + (int64_t)obs[(tb + r) * K + k]];
                    v[s2] *= m.normalize;
                    if (lane + 32 * s2 < N) vm = fmax(vm, v[s2]);
                }
                const double M = warp_max(vm);
#pragma unroll
                for (int s2 = 0; s2 < NS; ++s2) {
                    float d = M > -INFINITY ? (float)(v[s2] - M) : 0.f;
                    if (RATIO) d *= (float)ratios[tb + r];
                    float bl = __expf(d);
                    if (lane + 32 * s2 >= N) { d = 0.f; bl = 0.f; }
                    if (elog) elog[(tb + r) * NP + lane + 32 * s2] = d;
                    if (blin) blin[(tb + r) * NP + lane + 32 * s2] = bl;
                }
                if (lane == r) { mf_keep = 0.f; csum = M; }
            }
        }
        // ---- row maxima of the batch, one lane per row
        if (lane < rows) {
            double M = (double)mf_keep + csum;
            if (RATIO) M *= ratios[tb + lane];
            rowmax[tb + lane] = M;
            if (!(M > TEHMM_MINDBL) && seq_flag) emg_flag_row(tb + lane, seq_off, nseq, seq_flag);   // _emission.pyx:73-80
        }
    }
}

template <typename OBS, int GT, bool RATIO, int NS>
static cudaError_t launch_em4_3(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                                const double *ratios, float *elog, float *blin, double *rowmax,
                                int *seq_flag, int sms, int64_t row0, int64_t row1)
{
    const size_t smem = em4_smem_bytes(m);
    cudaError_t e = cudaFuncSetAttribute(emission_merged4_kernel<OBS, GT, RATIO, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t need = ((row1 - row0 + EM4_ROWS - 1) / EM4_ROWS + EM4_WARPS - 1) / EM4_WARPS;
    if (need < 1) need = 1;
    const int grid = (int)(need < sms ? need : sms);
    emission_merged4_kernel<OBS, GT, RATIO, NS><<<grid, EM4_WARPS * 32, smem, st>>>(m, (const OBS *)b.obs, row1, ratios, elog, blin,
                                                                                rowmax, seq_flag, b.seq_off, b.nseq, row0);
    return cudaGetLastError();
}

template <typename OBS>
static cudaError_t launch_em4(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                              const double *ratios, float *elog, float *blin, double *rowmax,
                              int *seq_flag, int sms, int64_t row0, int64_t row1)
{
#define EM4_GO1(GT, NS_) (ratios ? launch_em4_3<OBS, GT, true, NS_>(st, m, b, ratios, elog, blin, rowmax, seq_flag, sms, row0, row1) \
                                 : launch_em4_3<OBS, GT, false, NS_>(st, m, b, ratios, elog, blin, rowmax, seq_flag, sms, row0, row1))
#define EM4_GO(GT) (m.NS == 1 ? EM4_GO1(GT, 1) : EM4_GO1(GT, 2))
    switch (m.G) {
    case 1: return EM4_GO(1);
    case 2: return EM4_GO(2);
    case 3: return EM4_GO(3);
    case 4: return EM4_GO(4);
    case 5: return EM4_GO(5);
    case 6: return EM4_GO(6);
    case 7: return EM4_GO(7);
    default: return EM4_GO(8);
    }
#undef EM4_GO
#undef EM4_GO1
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
// whether rows of the batch can be run in pieces (tehmm_run_emission_rows): the four-rows-per-pass kernel only
bool tehmm_emission_rows_ok(const TehmmModelDev &m, int prec, const double *ratios)
{
    (void)ratios;
    return prec == TEHMM_F32 && m.G > 0 && m.G <= 8 && m.LD == 32 * m.NS && em4_smem_bytes(m) <= 227 * 1024;
}

// rows [row0, row1) of the batch; pieces must come in increasing order, the first one starting at
// row 0 (clears the per-sequence flags) and the last one ending at b.total (applies the
// _emission.pyx:59,73-80 fix-up).  Anything but [0, total) needs tehmm_emission_rows_ok().
int tehmm_launch_emission(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b, int prec,
                          const double *ratios, void *elog, void *blin, double *rowmax,
                          double *frame, int *seq_flag, int sms, cudaError_t *err, int64_t row0, int64_t row1)
{
    cudaError_t e = cudaSuccess;
    int launched = 0;
    if (row0 == 0) e = cudaMemsetAsync(seq_flag, 0, sizeof(int) * (size_t)b.nseq, st);
    if (e == cudaSuccess) {
        if (frame) e = launch_em_obs<float, 1>(st, m, b, ratios, nullptr, nullptr, nullptr, frame, seq_flag, sms);
        else if (prec == TEHMM_F32 && m.G > 0 && emg_smem_bytes(m) <= 227 * 1024) {
            // out-of-range symbols in the slow branch index the dense table: obs must be < S there too,
            // exactly the generic kernel's contract
            if (m.LD == 32 * m.NS && m.G <= 8 && em4_smem_bytes(m) <= 227 * 1024) {
                if (b.obs_bytes == 1) e = launch_em4<uint8_t>(st, m, b, ratios, (float *)elog, (float *)blin, rowmax, seq_flag, sms, row0, row1);
                else if (b.obs_bytes == 2) e = launch_em4<uint16_t>(st, m, b, ratios, (float *)elog, (float *)blin, rowmax, seq_flag, sms, row0, row1);
                else e = launch_em4<int32_t>(st, m, b, ratios, (float *)elog, (float *)blin, rowmax, seq_flag, sms, row0, row1);
            }
            else if (b.obs_bytes == 1) e = launch_emg<uint8_t>(st, m, b, ratios, (float *)elog, (float *)blin, rowmax, seq_flag, sms);
            else if (b.obs_bytes == 2) e = launch_emg<uint16_t>(st, m, b, ratios, (float *)elog, (float *)blin, rowmax, seq_flag, sms);
            else e = launch_emg<int32_t>(st, m, b, ratios, (float *)elog, (float *)blin, rowmax, seq_flag, sms);
        }
        else if (prec == TEHMM_F32) e = launch_em_obs<float, 0>(st, m, b, ratios, (float *)elog, (float *)blin, rowmax, nullptr, seq_flag, sms);
        else e = launch_em_obs<double, 0>(st, m, b, ratios, (double *)elog, (double *)blin, rowmax, nullptr, seq_flag, sms);
    }
    launched = 1;
    if (e == cudaSuccess && row1 == b.total) {
        launched = 2;
        int warps = 4;
        int grid = (int)((b.nseq + warps - 1) / warps);
        if (frame || prec == TEHMM_F32)
            emission_fix_kernel<float><<<grid, warps * 32, 0, st>>>(m.N, m.LD, seq_flag, b.seq_off, b.nseq, (float *)elog, (float *)blin, rowmax, frame);
        else
            emission_fix_kernel<double><<<grid, warps * 32, 0, st>>>(m.N, m.LD, seq_flag, b.seq_off, b.nseq, (double *)elog, (double *)blin, rowmax, frame);
        e = cudaGetLastError();
    }
    *err = e;
    return e == cudaSuccess ? launched : -1;
}
