// Warp-level building blocks for the chunked parallel-in-time trellis kernels.
// Internal header.
//
// Mapping used by every scan kernel: ONE WARP PER CHUNK of the time axis,
// lane l owns states j = l + 32*s (s < NS), the N x N transition matrix lives
// in registers (column j for forward/Viterbi, row i for backward), and the
// state vector is broadcast through a per-warp shared-memory line that every
// lane reads back with 128-bit loads.
#pragma once
#include "common.cuh"

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    float v[4];
    __device__ __forceinline__ void load(const float *p)
    {
        float4 q = *reinterpret_cast<const float4 *>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
};
template <> struct Vec4<double> {
    double v[4];
    __device__ __forceinline__ void load(const double *p)
    {
        double2 a = *reinterpret_cast<const double2 *>(p);
        double2 b = *reinterpret_cast<const double2 *>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};

template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TEHMM_FULL, v, o);
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T w = __shfl_xor_sync(TEHMM_FULL, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// y[s] = sum_i xs[i] * c[s][i]   (sum-product semiring)
template <typename T, int NS>
__device__ __forceinline__ void matvec_sum(const T *xs, const T (&c)[NS][32 * NS], T (&y)[NS])
{
    constexpr int NP = 32 * NS;
    T acc[NS][4];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = (T)0;
#pragma unroll
    for (int i = 0; i < NP; i += 4) {
        Vec4<T> x;
        x.load(xs + i);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            acc[s][0] = fma(x.v[0], c[s][i + 0], acc[s][0]);
            acc[s][1] = fma(x.v[1], c[s][i + 1], acc[s][1]);
            acc[s][2] = fma(x.v[2], c[s][i + 2], acc[s][2]);
            acc[s][3] = fma(x.v[3], c[s][i + 3], acc[s][3]);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) y[s] = (acc[s][0] + acc[s][1]) + (acc[s][2] + acc[s][3]);
}

// Same, and keeps the broadcast vector in registers for a second use.
template <typename T, int NS>
__device__ __forceinline__ void matvec_sum_keep(const T *xs, const T (&c)[NS][32 * NS],
                                                T (&y)[NS], T (&xv)[32 * NS])
{
    constexpr int NP = 32 * NS;
    T acc[NS][4];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = (T)0;
#pragma unroll
    for (int i = 0; i < NP; i += 4) {
        Vec4<T> x;
        x.load(xs + i);
        xv[i] = x.v[0]; xv[i + 1] = x.v[1]; xv[i + 2] = x.v[2]; xv[i + 3] = x.v[3];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            acc[s][0] = fma(x.v[0], c[s][i + 0], acc[s][0]);
            acc[s][1] = fma(x.v[1], c[s][i + 1], acc[s][1]);
            acc[s][2] = fma(x.v[2], c[s][i + 2], acc[s][2]);
            acc[s][3] = fma(x.v[3], c[s][i + 3], acc[s][3]);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) y[s] = (acc[s][0] + acc[s][1]) + (acc[s][2] + acc[s][3]);
}

// (max,+) semiring with the reference's tie rule: strict '>' scanning the
// from-state upward, so the lowest from-state wins (_hmm.pyx:232-247).
// extra0[s] is added to the from-state-0 candidate only (segment quirk).
template <typename T, int NS>
__device__ __forceinline__ void matvec_max(const T *xs, const T (&c)[NS][32 * NS],
                                           const T (&extra0)[NS], T (&y)[NS], int (&arg)[NS])
{
    constexpr int NP = 32 * NS;
#pragma unroll
    for (int i = 0; i < NP; i += 4) {
        Vec4<T> x;
        x.load(xs + i);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                T cand = x.v[q] + c[s][i + q];
                if (i + q == 0) {
                    y[s] = cand + extra0[s];
                    arg[s] = 0;
                } else if (cand > y[s]) {
                    y[s] = cand;
                    arg[s] = i + q;
                }
            }
        }
    }
}

// Scale a non-negative vector by a power of two so that its largest element
// lies in [1,2).  Exact (no rounding), hence independent of scaling history.
// Returns the exponent taken out: raw = canonical * 2^shift.
template <typename T, int NS> __device__ __forceinline__ int canonicalise(T (&x)[NS])
{
    T m = x[0];
#pragma unroll
    for (int s = 1; s < NS; ++s) m = x[s] > m ? x[s] : m;
    unsigned mb = __reduce_max_sync(TEHMM_FULL, TehmmNum<T>::order_bits(m));
    int e = sizeof(T) == 4 ? (int)(mb >> 23) : (int)(mb >> 20);
    if (e == 0) {
        // all zero (impossible data) or subnormal: lift by 2^64 and let the
        // next step finish the job
        if (!__any_sync(TEHMM_FULL, m > (T)0)) return 0;
        T up = TehmmNum<T>::inv_scale(TehmmNum<T>::BIAS - 64);
#pragma unroll
        for (int s = 0; s < NS; ++s) x[s] *= up;
        return -64;
    }
    if (e > TehmmNum<T>::EMAX) return 0;   // inf / nan: leave alone
    T sc = TehmmNum<T>::inv_scale(e);
#pragma unroll
    for (int s = 0; s < NS; ++s) x[s] *= sc;
    return e - TehmmNum<T>::BIAS;
}

// warp-wide maximum of arbitrary-sign values
__device__ __forceinline__ float warp_max_any(float v)
{
    unsigned b = __float_as_uint(v);
    unsigned key = b ^ ((unsigned)((int)b >> 31) | 0x80000000u);   // order-preserving
    key = __reduce_max_sync(TEHMM_FULL, key);
    b = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
    return __uint_as_float(b);
}
__device__ __forceinline__ double warp_max_any(double v) { return warp_max(v); }

// Subtract the maximum from a vector of log values (max becomes 0).
template <typename T, int NS> __device__ __forceinline__ void log_normalise(T (&d)[NS])
{
    T m = d[0];
#pragma unroll
    for (int s = 1; s < NS; ++s) m = d[s] > m ? d[s] : m;
    m = warp_max_any(m);
    if (m > (T)-INFINITY && m < (T)INFINITY) {
#pragma unroll
        for (int s = 0; s < NS; ++s) d[s] -= m;
    }
}
