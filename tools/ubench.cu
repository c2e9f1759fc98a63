// Micro-benchmarks of the instruction rates the scan kernels depend on (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32_k4(float (&d)[4], const uint32_t (&a)[2], const uint32_t (&b)[1])
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b[0]));
}

template <int KIND, int NACC>
__global__ void k_mma(float *out, long long *cyc)
{
    float d[NACC][4];
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 7 + i;
    for (int i = 0; i < 2; ++i) b[i] = threadIdx.x * 3 + i;
    for (int q = 0; q < NACC; ++q)
        for (int i = 0; i < 4; ++i) d[q][i] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            if (KIND == 0) mma_tf32(d[q], a, b);
            else if (KIND == 1) mma_bf16(d[q], a, b);
            else { uint32_t a2[2] = {a[0], a[1]}; uint32_t b1[1] = {b[0]}; mma_tf32_k4(d[q], a2, b1); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int q = 0; q < NACC; ++q)
        for (int i = 0; i < 4; ++i) s += d[q][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// KIND: 0 FFMA2, 1 FADD2, 2 FMNMX3 (fmaxf(fmaxf)), 3 FMNMX, 4 FFMA, 5 EX2, 6 FADD2+FMNMX3 mixed, 7 FMUL2
template <int KIND, int NACC>
__global__ void k_alu(float *out, long long *cyc, float seed)
{
    float2 x[NACC];
    float2 c = make_float2(seed, seed * 0.5f), e = make_float2(seed * 0.25f, 1.0f);
    for (int q = 0; q < NACC; ++q) x[q] = make_float2(threadIdx.x + q, threadIdx.x - q);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            if (KIND == 0) {
                asm volatile("{.reg .b64 a,b,c,d; mov.b64 a,{%0,%1}; mov.b64 b,{%2,%3}; mov.b64 c,{%4,%5}; fma.rn.f32x2 d,a,b,c; mov.b64 {%0,%1},d;}"
                             : "+f"(x[q].x), "+f"(x[q].y) : "f"(c.x), "f"(c.y), "f"(e.x), "f"(e.y));
            } else if (KIND == 1) {
                asm volatile("{.reg .b64 a,b,d; mov.b64 a,{%0,%1}; mov.b64 b,{%2,%3}; add.rn.f32x2 d,a,b; mov.b64 {%0,%1},d;}"
                             : "+f"(x[q].x), "+f"(x[q].y) : "f"(c.x), "f"(c.y));
            } else if (KIND == 2) {
                x[q].x = fmaxf(fmaxf(x[q].x, c.x), x[q].y + 0.f * e.x);
                asm volatile("" : "+f"(x[q].x));
            } else if (KIND == 3) {
                x[q].x = fmaxf(x[q].x, c.x);
                asm volatile("" : "+f"(x[q].x));
            } else if (KIND == 4) {
                x[q].x = fmaf(x[q].x, c.x, e.x);
                asm volatile("" : "+f"(x[q].x));
            } else if (KIND == 5) {
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[q].x));
            } else if (KIND == 6) {
                float2 s;
                asm volatile("{.reg .b64 a,b,d; mov.b64 a,{%2,%3}; mov.b64 b,{%4,%5}; add.rn.f32x2 d,a,b; mov.b64 {%0,%1},d;}"
                             : "=f"(s.x), "=f"(s.y) : "f"(x[q].y), "f"(e.x), "f"(c.x), "f"(c.y));
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[q].x) : "f"(s.x), "f"(s.y));
            } else if (KIND == 7) {
                asm volatile("{.reg .b64 a,b,d; mov.b64 a,{%0,%1}; mov.b64 b,{%2,%3}; mul.rn.f32x2 d,a,b; mov.b64 {%0,%1},d;}"
                             : "+f"(x[q].x), "+f"(x[q].y) : "f"(c.x), "f"(c.y));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int q = 0; q < NACC; ++q) s += x[q].x + x[q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// LDS rates: W = bytes per lane per load (4, 8, 16); broadcast pattern or lane-strided
template <int W, int BCAST>
__global__ void k_lds(float *out, long long *cyc)
{
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    float acc = 0;
    unsigned base = BCAST ? (threadIdx.x / 32) * 64 : threadIdx.x * (W / 4);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            unsigned idx = (base + q * 1024 + (it & 3) * 4) & 8191u;
            if (W == 16) { float4 v = *reinterpret_cast<const float4 *>(&sm[idx & ~3u]); acc += v.x + v.w; }
            else if (W == 8) { float2 v = *reinterpret_cast<const float2 *>(&sm[idx & ~1u]); acc += v.x + v.y; }
            else acc += sm[idx];
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void report(const char *name, F launch, int threads, double ops_per_iter_per_warp)
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 4 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 1024 * sizeof(long long));
    launch(out, cyc);
    cudaDeviceSynchronize();
    launch(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    int warps_per_smsp = threads / 32 / 4;
    double instr_per_clk_smsp = (double)ITERS * ops_per_iter_per_warp * warps_per_smsp / (double)mx;
    printf("%-34s thr=%4d cyc=%9lld  warp-instr/clk/SMSP=%.3f  (%s)\n", name, threads, mx, instr_per_clk_smsp, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
#define RUN_MMA(K, A, T) report("mma kind" #K " nacc" #A, [](float *o, long long *c) { k_mma<K, A><<<148, T>>>(o, c); }, T, A)
    RUN_MMA(0, 1, 128); RUN_MMA(0, 2, 128); RUN_MMA(0, 4, 128); RUN_MMA(0, 4, 256); RUN_MMA(0, 4, 384);
    RUN_MMA(0, 8, 128); RUN_MMA(0, 8, 256); RUN_MMA(0, 8, 512); RUN_MMA(0, 16, 256); RUN_MMA(0, 2, 1024);
    RUN_MMA(1, 8, 128); RUN_MMA(1, 8, 256); RUN_MMA(1, 8, 512);
    RUN_MMA(2, 8, 256); RUN_MMA(2, 8, 512);
#define RUN_ALU(K, A, T, nm) report(nm, [](float *o, long long *c) { k_alu<K, A><<<148, T>>>(o, c, 1.0001f); }, T, (K == 6 ? 2 * A : A))
    RUN_ALU(0, 8, 256, "FFMA2"); RUN_ALU(0, 8, 512, "FFMA2"); RUN_ALU(0, 8, 1024, "FFMA2");
    RUN_ALU(1, 8, 512, "FADD2"); RUN_ALU(7, 8, 512, "FMUL2");
    RUN_ALU(2, 8, 512, "FMNMX3?"); RUN_ALU(3, 8, 512, "FMNMX"); RUN_ALU(4, 8, 512, "FFMA");
    RUN_ALU(5, 8, 512, "EX2"); RUN_ALU(6, 8, 512, "FADD2+FMNMX3 (2 instr)"); RUN_ALU(6, 8, 1024, "FADD2+FMNMX3 (2 instr)");
#define RUN_LDS(W, B, T, nm) report(nm, [](float *o, long long *c) { k_lds<W, B><<<148, T>>>(o, c); }, T, 8)
    RUN_LDS(16, 1, 512, "LDS.128 bcast"); RUN_LDS(16, 0, 512, "LDS.128 strided");
    RUN_LDS(8, 1, 512, "LDS.64 bcast"); RUN_LDS(8, 0, 512, "LDS.64 strided");
    RUN_LDS(4, 1, 512, "LDS.32 bcast"); RUN_LDS(4, 0, 512, "LDS.32 strided");
    return 0;
}
