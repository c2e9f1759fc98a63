// Posterior-weighted per-track emission histograms for Baum-Welch
// (emission.py:221-241 -> _emission.pyx:171-190):
//     stats[k][j][obs[t][k]] += post[t][j] (* ratio[t])
//
// One CTA per contiguous slice of time, ONE WARP PER TRACK, lane = state.  The
// CTA keeps a private float64 histogram [symbol-row][state] in shared memory
// (same compact layout as the emission table), so warps never collide (they
// own disjoint rows) and no atomics are needed.  Annotation tracks are
// run-length structured, so each lane accumulates in a register while the
// track's symbol does not change and touches shared memory only on a change.
// CTA histograms go to global memory and are summed in fixed order (float64),
// which makes the result deterministic.
//
// Algorithmic HBM bytes per step: 4N (posteriors) + K (symbols).
#include "scan.cuh"
#include <algorithm>

#define ST_TILE 32   // time steps staged per CTA iteration

template <typename T, typename OBS>
__global__ void __launch_bounds__(1024)
emission_stats_kernel(TehmmModelDev m, const OBS *__restrict__ obs, int64_t total,
                      const T *__restrict__ post, const double *__restrict__ ratios,
                      double *__restrict__ part, double *__restrict__ dense_stats,
                      int64_t steps_per_cta, int track_base, int statS)
{
    extern __shared__ __align__(16) unsigned char st_smem[];
    const int N = m.N, K = m.K;
    const int64_t cells = (int64_t)m.tab_rows * N;
    double *hist = reinterpret_cast<double *>(st_smem);
    T *post_s = reinterpret_cast<T *>(hist + cells);                  // [ST_TILE][N]
    double *ratio_s = reinterpret_cast<double *>(post_s + ST_TILE * N + (ST_TILE * N & 1));
    int *sym_s = reinterpret_cast<int *>(ratio_s + ST_TILE);          // [ST_TILE][K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t e = threadIdx.x; e < cells; e += blockDim.x) hist[e] = 0.0;

    const int k = track_base + warp;
    const bool active = k < K;
    const int off = active ? m.tab_off[k] : 0, nsym = active ? m.track_nsym[k] : 0;
    const int64_t ta = (int64_t)blockIdx.x * steps_per_cta;
    const int64_t tz = min(total, ta + steps_per_cta);
    int cur = -1;
    double acc0 = 0.0, acc1 = 0.0;
    auto flush = [&]() {
        if (cur < 0) return;
        if (cur < nsym) {
            if (lane < N) hist[(int64_t)(off + cur) * N + lane] += acc0;
            if (lane + 32 < N) hist[(int64_t)(off + cur) * N + lane + 32] += acc1;
        } else {   // symbol outside the compact table: dense layout, rare
            if (lane < N) atomicAdd(&dense_stats[((int64_t)k * N + lane) * statS + cur], acc0);
            if (lane + 32 < N) atomicAdd(&dense_stats[((int64_t)k * N + lane + 32) * statS + cur], acc1);
        }
    };
    for (int64_t tb = ta; tb < tz; tb += ST_TILE) {
        const int rows = (int)min((int64_t)ST_TILE, tz - tb);
        __syncthreads();
        for (int e = threadIdx.x; e < rows * N; e += blockDim.x) post_s[e] = post[(tb + e / N) * m.LD + e % N];
        for (int e = threadIdx.x; e < rows * K; e += blockDim.x) sym_s[e] = (int)obs[tb * K + e];
        if (ratios)
            for (int e = threadIdx.x; e < rows; e += blockDim.x) ratio_s[e] = ratios[tb + e];
        __syncthreads();
        if (!active) continue;
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
            const int sym = sym_s[r * K + k];
            double g0 = lane < N ? (double)post_s[r * N + lane] : 0.0;
            double g1 = lane + 32 < N ? (double)post_s[r * N + lane + 32] : 0.0;
            if (ratios) { const double rr = ratio_s[r]; g0 *= rr; g1 *= rr; }
            if (sym != cur) {
                flush();
                cur = sym;
                acc0 = 0.0;
                acc1 = 0.0;
            }
            acc0 += g0;
            acc1 += g1;
        }
    }
    if (active) flush();
    __syncthreads();
    double *dst = part + (int64_t)blockIdx.x * cells;
    for (int64_t e = threadIdx.x; e < cells; e += blockDim.x) dst[e] = hist[e];
}

// fp32 production path (N <= 32, lattice stride 32, K <= 32, no segment ratios):
// every warp takes batches of 16 time steps, lane = state, and adds its posterior
// row into a CTA-wide histogram with NATIVE shared-memory atomics.  Floating-point
// atomics on shared memory are compare-and-swap loops on sm_100a (ATOMS.CAST.SPIN),
// 64-bit integer ones too; 32-bit integer adds are native (ATOMS.ADD).  So a
// posterior p in [0,1] is added as the 40-bit fixed-point number round(p * 2^40),
// split into a 20-bit low limb and a high limb of at most 21 bits, each into its own
// uint32 bin; the bins are flushed into the CTA's float64 partial every SA_FLUSH *
// SA_TILE = 1536 steps (1536 * 2^21 < 2^32).  Quantisation 2^-41 per addition, and -- integer
// addition being associative -- the result is bit-reproducible whatever the order
// in which warps get to the bins.  All 32 warps are busy (the one-warp-per-track
// kernel above keeps K of them busy and chains read-modify-writes).
#define SA_WARPS 32
#define SA_ROWS 16
#define SA_TILE (SA_WARPS * SA_ROWS)
#define SA_FLUSH 3       // tiles between flushes: 3 * 512 * 2^21 < 2^32

template <typename OBS>
__global__ void __launch_bounds__(SA_WARPS * 32, 1)
emission_stats_atomic_kernel(TehmmModelDev m, const OBS *__restrict__ obs, int64_t total,
                             const float *__restrict__ post, double *__restrict__ part,
                             double *__restrict__ dense_stats, int64_t steps_per_cta, int statS)
{
    extern __shared__ __align__(16) unsigned char st_smem[];
    const int N = m.N, K = m.K, KP = (K + 3) & ~3;
    unsigned *hist = reinterpret_cast<unsigned *>(st_smem);                    // [tab_rows][2 limbs][32]
    int32_t *koff = reinterpret_cast<int32_t *>(hist + (size_t)m.tab_rows * 64);  // [K] first row, [K] widths
    int32_t *offs_all = koff + 2 * K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *offs = offs_all + (size_t)warp * SA_ROWS * KP;
    const int64_t cells = (int64_t)m.tab_rows * N;
    for (int e = threadIdx.x; e < m.tab_rows * 64; e += blockDim.x) hist[e] = 0u;
    for (int e = threadIdx.x; e < K; e += blockDim.x) { koff[e] = m.tab_off[e]; koff[K + e] = m.track_nsym[e]; }
    double *dst = part + (int64_t)blockIdx.x * cells;
    for (int64_t e = threadIdx.x; e < cells; e += blockDim.x) dst[e] = 0.0;
    __syncthreads();
    const uint32_t hist_lane = (uint32_t)__cvta_generic_to_shared(hist) + (uint32_t)lane * 4u;

    const int64_t ta = (int64_t)blockIdx.x * steps_per_cta;
    const int64_t tz = min(total, ta + steps_per_cta);
    int since = 0;
    for (int64_t tile = ta; tile < tz; tile += SA_TILE) {
        const int64_t tb = tile + (int64_t)warp * SA_ROWS;
        const int rows = (int)max((int64_t)0, min((int64_t)SA_ROWS, tz - tb));
        __syncwarp();                     // the previous tile's reads of offs are done
        // the batch's posterior rows (coalesced, all in flight) and shared-memory offsets of its symbols
        float p[SA_ROWS];
#pragma unroll
        for (int r = 0; r < SA_ROWS; ++r) p[r] = r < rows ? post[(tb + r) * 32 + lane] : 0.f;
        for (int e = lane; e < rows * K; e += 32) {
            const int r = e / K, k = e - r * K;
            const int sym = (int)obs[tb * K + e];
            offs[r * KP + k] = sym < koff[K + k] ? (koff[k] + sym) * 256 : -(sym + 1);
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < SA_ROWS; ++r) {
            if (r < rows) {
                // a posterior is in [0,1] up to rounding; clamp so the high limb stays below 2^19
                const unsigned long long q = __float2ull_rn(fminf(p[r], 2.f) * 1099511627776.f);   // * 2^40
                const unsigned lo = (unsigned)q & 0xfffffu, hi = (unsigned)(q >> 20);
                for (int k = 0; k < K; ++k) {
                    const int o = offs[r * KP + k];
                    if (o >= 0) {
                        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hist_lane + (uint32_t)o), "r"(lo) : "memory");
                        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hist_lane + (uint32_t)o + 128u), "r"(hi) : "memory");
                    } else if (lane < N) {  // symbol outside the compact table: dense layout, rare
                        atomicAdd(&dense_stats[((int64_t)k * N + lane) * statS + (-o - 1)], (double)p[r]);
                    }
                }
            }
        }
        // flush the fixed-point histogram into the CTA's float64 partial (warp per row) every
        // SA_FLUSH tiles: SA_FLUSH * SA_TILE additions of at most 2^21 / 2^20 fit the 32-bit bins
        if (++since == SA_FLUSH || tile + SA_TILE >= tz) {
            since = 0;
            __syncthreads();
            for (int row = warp; row < m.tab_rows; row += SA_WARPS) {
                const unsigned lo = hist[row * 64 + lane], hi = hist[row * 64 + 32 + lane];
                if ((lo | hi) && lane < N) {
                    dst[(int64_t)row * N + lane] += ((double)hi * 1048576.0 + (double)lo) * 9.094947017729282e-13;    // 2^20, 2^-40
                    hist[row * 64 + lane] = 0u;
                    hist[row * 64 + 32 + lane] = 0u;
                }
            }
            __syncthreads();
        }
    }
}

// *max_bits = bit pattern of max_t (float)ratios[t] (ratios are positive: patterns order like values);
// zeroed by the launcher.  The consumer rounds it up to a power of two (ratio_cap_of).
__global__ void ratio_max_kernel(const double *__restrict__ ratios, int64_t total, unsigned *__restrict__ max_bits)
{
    float mx = 0.f;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        mx = fmaxf(mx, (float)ratios[t]);
    unsigned b = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(mx, 0.f)));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits, b);
}
__device__ __forceinline__ float ratio_cap_of(const float *cap_dev)
{
    // (float)ratio rounds to nearest, so a ratio just below a power of two may round up to it: one binade of slack
    const unsigned b = __float_as_uint(fmaxf(*cap_dev, 1.f));
    return __uint_as_float(((b >> 23) + 1u) << 23);
}

// Same, over MERGED tracks (TehmmModelDev::sgdesc): a group of up to four tracks has one
// histogram row per combination of its tracks' symbols, so a time step costs two reductions
// per GROUP instead of two per track (10 tracks -> 5 groups at the bench shape: the kernel is
// bound by the shared-memory atomic rate, ~1 per 3.3 cycles per SM).  The per-track histograms
// are the marginals of the merged ones, taken in emission_stats_unmerge_kernel (exact: the
// merged counts are sums of the same fixed-point posteriors).
template <typename OBS, int SGT>
__global__ void __launch_bounds__(SA_WARPS * 32, 1)
emission_stats_merged_kernel(TehmmModelDev m, const OBS *__restrict__ obs, int64_t total,
                             const float *__restrict__ post, double *__restrict__ part,
                             double *__restrict__ dense_stats, int64_t steps_per_cta, int statS,
                             const double *__restrict__ ratios, const float *__restrict__ cap_dev)
{
    // ratios != nullptr (fastAccumulateStats with segRatios, _emission.pyx:146-234): a row's posteriors are
    // weighted by its segment ratio.  The fixed-point bins hold weight / cap with cap = the power of two at or
    // above the largest ratio of the batch (ratio_cap_kernel), so they keep their 41-bit headroom; the flush
    // multiplies it back (a power of two: exact).
    const float cap = ratios ? ratio_cap_of(cap_dev) : 1.f, inv_cap = 1.f / cap;
    extern __shared__ __align__(16) unsigned char st_smem[];
    const int N = m.N, K = m.K;
    unsigned *hist = reinterpret_cast<unsigned *>(st_smem);                        // [srows][2 limbs][32]
    int32_t *gd_s = reinterpret_cast<int32_t *>(hist + (size_t)m.srows * 64);      // [SGT][TEHMM_GDESC]
    int32_t *nsym_s = gd_s + 8 * TEHMM_GDESC;                                      // [K]
    int32_t *offs_all = nsym_s + ((K + 3) & ~3);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *offs = offs_all + (size_t)warp * SA_ROWS * 8;
    const int64_t cells = (int64_t)m.srows * N;
    for (int e = threadIdx.x; e < m.srows * 64; e += blockDim.x) hist[e] = 0u;
    for (int e = threadIdx.x; e < SGT * TEHMM_GDESC; e += blockDim.x) gd_s[e] = m.sgdesc[e];
    for (int e = threadIdx.x; e < K; e += blockDim.x) nsym_s[e] = m.track_nsym[e];
    double *dst = part + (int64_t)blockIdx.x * cells;
    for (int64_t e = threadIdx.x; e < cells; e += blockDim.x) dst[e] = 0.0;
    __syncthreads();
    const uint32_t hist_lane = (uint32_t)__cvta_generic_to_shared(hist) + (uint32_t)lane * 4u;

    const int64_t ta = (int64_t)blockIdx.x * steps_per_cta;
    const int64_t tz = min(total, ta + steps_per_cta);
    int since = 0;
    for (int64_t tile = ta; tile < tz; tile += SA_TILE) {
        const int64_t tb = tile + (int64_t)warp * SA_ROWS;
        const int rows = (int)max((int64_t)0, min((int64_t)SA_ROWS, tz - tb));
        __syncwarp();                     // the previous tile's reads of offs are done
        float p[SA_ROWS];
#pragma unroll
        for (int r = 0; r < SA_ROWS; ++r) p[r] = r < rows ? post[(tb + r) * 32 + lane] : 0.f;
        if (ratios) {
#pragma unroll
            for (int r = 0; r < SA_ROWS; ++r)
                if (r < rows) p[r] *= (float)ratios[tb + r] * inv_cap;
        }
        // byte offset of the merged histogram row of every (row, group); -1: a symbol outside its table
        for (int e = lane; e < rows * SGT; e += 32) {
            const int r = e / SGT, gq = e - r * SGT;
            const int32_t *d = gd_s + gq * TEHMM_GDESC;
            int idx = d[1];
            bool bad = false;
            for (int i = 0; i < d[0]; ++i) {
                const int k = d[2 + i];
                const int sym = (int)obs[(tb + r) * K + k];
                bad |= (unsigned)sym >= (unsigned)nsym_s[k];
                idx += sym * d[6 + i];
            }
            offs[r * 8 + gq] = bad ? -1 : idx * 256;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < SA_ROWS; ++r) {
            if (r < rows) {
                const unsigned long long q = __float2ull_rn(fminf(p[r], 2.f) * 1099511627776.f);   // * 2^40
                const unsigned lo = (unsigned)q & 0xfffffu, hi = (unsigned)(q >> 20);
#pragma unroll
                for (int gq = 0; gq < SGT; ++gq) {
                    const int o = offs[r * 8 + gq];
                    if (o >= 0) {
                        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hist_lane + (uint32_t)o), "r"(lo) : "memory");
                        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hist_lane + (uint32_t)o + 128u), "r"(hi) : "memory");
                    } else if (lane < N) {   // rare: every track of the group goes to the dense layout
                        const int32_t *d = gd_s + gq * TEHMM_GDESC;
                        for (int i = 0; i < d[0]; ++i) {
                            const int k = d[2 + i];
                            atomicAdd(&dense_stats[((int64_t)k * N + lane) * statS + (int)obs[(tb + r) * K + k]], (double)p[r] * (double)cap);
                        }
                    }
                }
            }
        }
        if (++since == SA_FLUSH || tile + SA_TILE >= tz) {
            since = 0;
            __syncthreads();
            for (int row = warp; row < m.srows; row += SA_WARPS) {
                const unsigned lo = hist[row * 64 + lane], hi = hist[row * 64 + 32 + lane];
                if ((lo | hi) && lane < N) {
                    dst[(int64_t)row * N + lane] += ((double)hi * 1048576.0 + (double)lo) * 9.094947017729282e-13 * (double)cap;    // 2^20, 2^-40
                    hist[row * 64 + lane] = 0u;
                    hist[row * 64 + 32 + lane] = 0u;
                }
            }
            __syncthreads();
        }
    }
}

// part[0][row][j] = sum over CTAs of part[cta][row][j], fixed order (float64)
__global__ void emission_stats_sum_parts_kernel(int64_t cells, double *__restrict__ part, int nparts)
{
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= cells) return;
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += part[(int64_t)p * cells + e];
    part[e] = acc;
}

// obs_stats[k][j][sym] += sum over the merged rows of k's group whose k-digit is sym (fixed
// order, float64): one thread per (compact row of track k, state j); H = the summed histogram.
__global__ void emission_stats_unmerge_kernel(TehmmModelDev m, const double *__restrict__ H,
                                              double *__restrict__ obs_stats, int statS)
{
    const int N = m.N;
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= (int64_t)m.tab_rows * N) return;
    const int row = (int)(e / N), j = (int)(e - (int64_t)row * N);
    int k = 0;
    while (k + 1 < m.K && m.tab_off[k + 1] <= row) ++k;
    const int sym = row - m.tab_off[k];
    if (sym >= statS) return;
    // k's group, its stride there and the group's size
    int gq = 0, pos = 0;
    for (int g2 = 0; g2 < m.SG; ++g2)
        for (int i = 0; i < m.sgdesc[g2 * TEHMM_GDESC]; ++i)
            if (m.sgdesc[g2 * TEHMM_GDESC + 2 + i] == k) { gq = g2; pos = i; }
    const int32_t *d = m.sgdesc + gq * TEHMM_GDESC;
    int64_t grows = 1;
    for (int i = 0; i < d[0]; ++i) grows *= m.track_nsym[d[2 + i]];
    const int64_t stride = d[6 + pos], nk = m.track_nsym[k];
    double acc = 0.0;
    for (int64_t hi = 0; hi < grows / (stride * nk); ++hi)
        for (int64_t lo = 0; lo < stride; ++lo) acc += H[(d[1] + (hi * nk + sym) * stride + lo) * N + j];
    obs_stats[((int64_t)k * N + j) * statS + sym] += acc;
}

// Slow path when the compact histogram does not fit shared memory: global atomics.
template <typename T, typename OBS>
__global__ void emission_stats_global_kernel(TehmmModelDev m, const OBS *__restrict__ obs,
                                             int64_t total, const T *__restrict__ post,
                                             const double *__restrict__ ratios,
                                             double *__restrict__ dense_stats, int statS)
{
    const int N = m.N, K = m.K;
    const int64_t cells = total * K;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < cells;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = e / K;
        const int k = (int)(e - t * K);
        const int sym = (int)obs[e];
        const double r = ratios ? ratios[t] : 1.0;
        for (int j = 0; j < N; ++j)
            atomicAdd(&dense_stats[((int64_t)k * N + j) * statS + sym], (double)post[t * m.LD + j] * r);
    }
}

// obs_stats[k][j][sym] += sum over CTAs of part[cta][off_k+sym][j]
__global__ void emission_stats_reduce_kernel(TehmmModelDev m, const double *__restrict__ part,
                                             int nparts, double *__restrict__ obs_stats, int statS)
{
    const int N = m.N;
    const int64_t cells = (int64_t)m.tab_rows * N;
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= cells) return;
    const int row = (int)(e / N), j = (int)(e - (int64_t)row * N);
    int k = 0;
    while (k + 1 < m.K && m.tab_off[k + 1] <= row) ++k;
    const int sym = row - m.tab_off[k];
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += part[(int64_t)p * cells + e];
    if (sym < statS) obs_stats[((int64_t)k * N + j) * statS + sym] += acc;
}

template <typename T, typename OBS>
static cudaError_t launch_stats(cudaStream_t st, const TehmmModelDev &m, const TehmmBatchDev &b,
                                const T *post, const double *ratios, double *obs_stats,
                                double *part, int nparts, int statS)
{
    const size_t smem = (size_t)m.tab_rows * m.N * sizeof(double) + (size_t)(ST_TILE * m.N + 1) * sizeof(T)
                        + ST_TILE * sizeof(double) + (size_t)ST_TILE * m.K * sizeof(int) + 16;
    if (nparts <= 0) {
        emission_stats_global_kernel<T, OBS><<<148 * 8, 256, 0, st>>>(m, (const OBS *)b.obs, b.total, post, ratios, obs_stats, statS);
        return cudaGetLastError();
    }
    if (sizeof(T) == 4 && m.LD == 32 && m.SG >= 1 && m.SG <= 8 && m.SG < m.K) {
        const size_t sm3 = (size_t)m.srows * 256 + (size_t)(8 * TEHMM_GDESC + ((m.K + 3) & ~3)) * 4 + (size_t)SA_WARPS * SA_ROWS * 8 * 4 + 16;
        if (sm3 <= 220 * 1024) {
            const int64_t per = ((b.total + nparts - 1) / nparts + SA_TILE - 1) / SA_TILE * SA_TILE;
            // the ratio cap lives behind the CTA partials (part has room for max(tab_rows, srows) * N doubles per CTA)
            float *cap_dev = reinterpret_cast<float *>(part + (int64_t)nparts * std::max(m.tab_rows, m.srows) * m.N);
            if (ratios) {
                cudaError_t ec = cudaMemsetAsync(cap_dev, 0, sizeof(float), st);
                if (ec != cudaSuccess) return ec;
                ratio_max_kernel<<<148 * 4, 256, 0, st>>>(ratios, b.total, reinterpret_cast<unsigned *>(cap_dev));
            }
#define ST_MERGED(G_) do { auto k3 = emission_stats_merged_kernel<OBS, G_>; \
                           cudaError_t e3 = cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3); \
                           if (e3 != cudaSuccess) return e3; \
                           k3<<<nparts, SA_WARPS * 32, sm3, st>>>(m, (const OBS *)b.obs, b.total, (const float *)post, part, obs_stats, per, statS, ratios, cap_dev); } while (0)
            switch (m.SG) {
            case 1: ST_MERGED(1); break;
            case 2: ST_MERGED(2); break;
            case 3: ST_MERGED(3); break;
            case 4: ST_MERGED(4); break;
            case 5: ST_MERGED(5); break;
            case 6: ST_MERGED(6); break;
            case 7: ST_MERGED(7); break;
            default: ST_MERGED(8); break;
            }
#undef ST_MERGED
            const int64_t cells = (int64_t)m.tab_rows * m.N;
            const int64_t mcells = (int64_t)m.srows * m.N;
            emission_stats_sum_parts_kernel<<<(int)((mcells + 127) / 128), 128, 0, st>>>(mcells, part, nparts);
            emission_stats_unmerge_kernel<<<(int)((cells + 127) / 128), 128, 0, st>>>(m, part, obs_stats, statS);
            return cudaGetLastError();
        }
    }
    if (sizeof(T) == 4 && m.LD == 32 && m.K <= 32 && !ratios) {
        const int KP = (m.K + 3) & ~3;
        const size_t sm2 = (size_t)m.tab_rows * 256 + (size_t)2 * m.K * 4 + (size_t)SA_WARPS * SA_ROWS * KP * 4 + 16;
        if (sm2 <= 220 * 1024) {
            const int64_t per = ((b.total + nparts - 1) / nparts + SA_TILE - 1) / SA_TILE * SA_TILE;
            auto k2 = emission_stats_atomic_kernel<OBS>;
            cudaError_t e2 = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
            if (e2 != cudaSuccess) return e2;
            k2<<<nparts, SA_WARPS * 32, sm2, st>>>(m, (const OBS *)b.obs, b.total, (const float *)post, part, obs_stats, per, statS);
            const int64_t cells = (int64_t)m.tab_rows * m.N;
            emission_stats_reduce_kernel<<<(int)((cells + 127) / 128), 128, 0, st>>>(m, part, nparts, obs_stats, statS);
            return cudaGetLastError();
        }
    }
    auto kern = emission_stats_kernel<T, OBS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t steps = (b.total + nparts - 1) / nparts;
    for (int base = 0; base < m.K; base += 32) {
        int warps = m.K - base < 32 ? m.K - base : 32;
        kern<<<nparts, warps * 32, smem, st>>>(m, (const OBS *)b.obs, b.total, post, ratios, part, obs_stats, steps, base, statS);
        const int64_t cells = (int64_t)m.tab_rows * m.N;
        emission_stats_reduce_kernel<<<(int)((cells + 127) / 128), 128, 0, st>>>(m, part, nparts, obs_stats, statS);
    }
    return cudaGetLastError();
}

// nparts = number of CTA-private histograms (0 = global-atomics slow path)
cudaError_t tehmm_launch_emission_stats(cudaStream_t st, const TehmmModelDev &m,
                                        const TehmmBatchDev &b, int prec, const void *post,
                                        const double *ratios, double *obs_stats, double *part,
                                        int nparts, int statS)
{
#define STATS_GO(TT)                                                                              \
    do {                                                                                          \
        if (b.obs_bytes == 1) return launch_stats<TT, uint8_t>(st, m, b, (const TT *)post, ratios, obs_stats, part, nparts, statS);   \
        if (b.obs_bytes == 2) return launch_stats<TT, uint16_t>(st, m, b, (const TT *)post, ratios, obs_stats, part, nparts, statS);  \
        return launch_stats<TT, int32_t>(st, m, b, (const TT *)post, ratios, obs_stats, part, nparts, statS);                         \
    } while (0)
    if (prec == TEHMM_F32) STATS_GO(float);
    STATS_GO(double);
#undef STATS_GO
}

size_t tehmm_stats_smem_bytes(int tab_rows, int N, int K, int prec)
{
    size_t ts = prec == TEHMM_F32 ? sizeof(float) : sizeof(double);
    return (size_t)tab_rows * N * sizeof(double) + (size_t)(ST_TILE * N + 1) * ts +
           ST_TILE * sizeof(double) + (size_t)ST_TILE * K * sizeof(int) + 16;
}
