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

// block-wide sum (blockDim.x a multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double part[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) part[w] = v;
    __syncthreads();
    v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
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

// ---- Blackwell packed fp32 pairs (FFMA2 / FADD2, sm_100+) and 3-input max (FMNMX3)
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi)
{
    return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo);
}
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Lane-owned slice of the (padded) transition matrix: column j (forward,
// Viterbi) or row i (backward) for each owned state.  float: packed pairs.
template <typename T, int NS> struct MatSlice {
    T c[NS][32 * NS];
    __device__ __forceinline__ void set(int s, int i, T v) { c[s][i] = v; }
};
template <int NS> struct MatSlice<float, NS> {
    u64 c[NS][16 * NS];
    float tmp;
    __device__ __forceinline__ void set(int s, int i, float v)
    {
        if (i & 1) c[s][i >> 1] = pk2(tmp, v); else tmp = v;
    }
};

// y[s] = sum_i xs[i] * M[s][i]
template <int NS>
__device__ __forceinline__ void matvec_sum(const float *xs, const MatSlice<float, NS> &M, float (&y)[NS])
{
    constexpr int NP = 32 * NS;
    u64 a[NS][4];
#pragma unroll
    for (int s = 0; s < NS; ++s) a[s][0] = a[s][1] = a[s][2] = a[s][3] = 0ull;
#pragma unroll
    for (int i = 0; i < NP; i += 8) {
        const ulonglong2 x0 = *reinterpret_cast<const ulonglong2 *>(xs + i);
        const ulonglong2 x1 = *reinterpret_cast<const ulonglong2 *>(xs + i + 4);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            a[s][0] = ffma2(x0.x, M.c[s][i / 2 + 0], a[s][0]);
            a[s][1] = ffma2(x0.y, M.c[s][i / 2 + 1], a[s][1]);
            a[s][2] = ffma2(x1.x, M.c[s][i / 2 + 2], a[s][2]);
            a[s][3] = ffma2(x1.y, M.c[s][i / 2 + 3], a[s][3]);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const u64 t = fadd2(fadd2(a[s][0], a[s][1]), fadd2(a[s][2], a[s][3]));
        y[s] = lo2(t) + hi2(t);
    }
}
template <int NS>
__device__ __forceinline__ void matvec_sum(const double *xs, const MatSlice<double, NS> &M, double (&y)[NS])
{
    matvec_sum<double, NS>(xs, M.c, y);
}

// same, keeping the broadcast vector in registers (packed for float) for a second use
template <int NS>
__device__ __forceinline__ void matvec_sum_keep(const float *xs, const MatSlice<float, NS> &M,
                                                float (&y)[NS], u64 (&xv)[16 * NS])
{
    constexpr int NP = 32 * NS;
    u64 a[NS][4];
#pragma unroll
    for (int s = 0; s < NS; ++s) a[s][0] = a[s][1] = a[s][2] = a[s][3] = 0ull;
#pragma unroll
    for (int i = 0; i < NP; i += 8) {
        const ulonglong2 x0 = *reinterpret_cast<const ulonglong2 *>(xs + i);
        const ulonglong2 x1 = *reinterpret_cast<const ulonglong2 *>(xs + i + 4);
        xv[i / 2] = x0.x; xv[i / 2 + 1] = x0.y; xv[i / 2 + 2] = x1.x; xv[i / 2 + 3] = x1.y;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            a[s][0] = ffma2(x0.x, M.c[s][i / 2 + 0], a[s][0]);
            a[s][1] = ffma2(x0.y, M.c[s][i / 2 + 1], a[s][1]);
            a[s][2] = ffma2(x1.x, M.c[s][i / 2 + 2], a[s][2]);
            a[s][3] = ffma2(x1.y, M.c[s][i / 2 + 3], a[s][3]);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const u64 t = fadd2(fadd2(a[s][0], a[s][1]), fadd2(a[s][2], a[s][3]));
        y[s] = lo2(t) + hi2(t);
    }
}

// m[s] = max_i (xs[i] + M[s][i])   ((max,+) semiring, value only)
template <int NS>
__device__ __forceinline__ void matvec_maxval(const float *xs, const MatSlice<float, NS> &M, float (&m)[NS])
{
    constexpr int NP = 32 * NS;
    float a[NS][2];
#pragma unroll
    for (int s = 0; s < NS; ++s) a[s][0] = a[s][1] = -INFINITY;
#pragma unroll
    for (int i = 0; i < NP; i += 4) {
        const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(xs + i);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const u64 c0 = fadd2(x.x, M.c[s][i / 2]), c1 = fadd2(x.y, M.c[s][i / 2 + 1]);
            a[s][0] = fmax3(a[s][0], lo2(c0), hi2(c0));
            a[s][1] = fmax3(a[s][1], lo2(c1), hi2(c1));
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) m[s] = fmaxf(a[s][0], a[s][1]);
}
template <int NS>
__device__ __forceinline__ void matvec_maxval(const double *xs, const MatSlice<double, NS> &M, double (&m)[NS])
{
    constexpr int NP = 32 * NS;
#pragma unroll
    for (int s = 0; s < NS; ++s) m[s] = -INFINITY;
#pragma unroll
    for (int i = 0; i < NP; i += 4) {
        Vec4<double> x;
        x.load(xs + i);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
#pragma unroll
            for (int q = 0; q < 4; ++q) m[s] = fmax(m[s], x.v[q] + M.c[s][i + q]);
        }
    }
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
// Returns the exponent taken out: raw = scaled * 2^shift.
template <typename T> struct ScaleShift { T sc; int shift; };
template <typename T> __device__ __noinline__ ScaleShift<T> canonicalise_rare(T m)
{
    // called warp-uniformly: the largest element is zero (impossible data),
    // subnormal, inf or nan
    ScaleShift<T> r;
    r.sc = (T)1;
    r.shift = 0;
    if (!__any_sync(TEHMM_FULL, m > (T)0)) return r;
    if (__any_sync(TEHMM_FULL, !(m < (T)INFINITY))) return r;
    // subnormal maximum: lift by 2^64 and let the next step finish the job
    r.shift = -64;
    r.sc = TehmmNum<T>::inv_scale(TehmmNum<T>::BIAS - 64);
    return r;
}
template <typename T, int NS> __device__ __forceinline__ int canonicalise(T (&x)[NS])
{
    T m = x[0];
#pragma unroll
    for (int s = 1; s < NS; ++s) m = x[s] > m ? x[s] : m;
    const unsigned mb = __reduce_max_sync(TEHMM_FULL, TehmmNum<T>::order_bits(m));
    const unsigned e = sizeof(T) == 4 ? (mb >> 23) : (mb >> 20);
    ScaleShift<T> r;
    if (__builtin_expect(e - 1u >= (unsigned)TehmmNum<T>::EMAX, 0)) {
        r = canonicalise_rare<T>(m);
    } else {
        r.sc = TehmmNum<T>::inv_scale((int)e);
        r.shift = (int)e - TehmmNum<T>::BIAS;
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) x[s] *= r.sc;
    return r.shift;
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
