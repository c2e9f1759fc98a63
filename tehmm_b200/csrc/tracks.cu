// Data formats either side of the trellis, on the device, so
// that the (T, K) symbol matrix the HMM reads is BUILT in HBM instead of travelling there:
//
//  * rasterisation: BED intervals -> one column of the (T, K) table
//    (trackIO.py:67-212 readBedData: a Python loop over every base of every interval);
//  * segmentation: variable-length segments from the columns of the table
//    (bin/segmentTracks.py:200-277 segmentTracks / isNewSegment: a Python loop over every base);
//  * compression: one row per segment, the row being the per-track MODE of the segment
//    (track.py:449-533,603-620 TrackTable.segment / interpolateSegments / setAverages /
//    compressSegments: scipy.stats.mode per segment per track);
//  * runSum (_track.pyx:13-25): exclusive running count of the zeros of a mask.
//
// All of it is byte / index work and bit-exact against the reference.  Interval parsing, value maps
// (strings -> categories) and file formats stay on the host (tehmm_b200/trackIO.py).
#include "common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

void tehmm_set_error(int code, const char *msg);
int tehmm_ctx_device(tehmm_ctx *c);
extern "C" uint64_t tehmm_ctx_stream(tehmm_ctx *c);

namespace {

int terr(int code, const char *msg) { tehmm_set_error(code, msg); return code; }
#define TCU(x)                                                                       \
    do {                                                                             \
        cudaError_t e__ = (x);                                                       \
        if (e__ != cudaSuccess) return terr(TEHMM_ECUDA, cudaGetErrorString(e__));   \
    } while (0)

struct TmpBuf {
    void *p = nullptr;
    ~TmpBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
    template <typename T> T *as() { return (T *)p; }
};

__device__ __forceinline__ void store_elem(void *table, int elem_bytes, int64_t idx, int v)
{
    if (elem_bytes == 1) ((uint8_t *)table)[idx] = (uint8_t)v;
    else if (elem_bytes == 2) ((uint16_t *)table)[idx] = (uint16_t)v;
    else ((int32_t *)table)[idx] = v;
}
__device__ __forceinline__ int load_elem(const void *table, int elem_bytes, int64_t idx)
{
    if (elem_bytes == 1) return ((const uint8_t *)table)[idx];
    if (elem_bytes == 2) return ((const uint16_t *)table)[idx];
    return ((const int32_t *)table)[idx];
}

// ---------------------------------------------------------------- rasterisation
// readBedData writes interval after interval, so where intervals overlap the LAST one in file order
// wins (trackIO.py:198-202).  owner[x] = index of the last interval covering base x.
__global__ void paint_owner_kernel(const int64_t *__restrict__ starts, const int64_t *__restrict__ ends,
                                   int64_t n, int *__restrict__ owner)
{
    // one warp per interval
    const int lane = threadIdx.x & 31;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += ((int64_t)gridDim.x * blockDim.x) >> 5)
        for (int64_t x = starts[i] + lane; x < ends[i]; x += 32) atomicMax(owner + x, (int)i);
}
// long intervals: one block per interval piece of PAINT_PIECE bases
#define PAINT_PIECE 8192
__global__ void paint_owner_long_kernel(const int64_t *__restrict__ starts, const int64_t *__restrict__ ends,
                                        const int64_t *__restrict__ piece_iv, const int64_t *__restrict__ piece_off,
                                        int64_t npieces, int *__restrict__ owner)
{
    for (int64_t p = blockIdx.x; p < npieces; p += gridDim.x) {
        const int64_t i = piece_iv[p];
        const int64_t a = starts[i] + piece_off[p], e = min(ends[i], a + (int64_t)PAINT_PIECE);
        for (int64_t x = a + threadIdx.x; x < e; x += blockDim.x) atomicMax(owner + x, (int)i);
    }
}
// data[oStart - start] = val0; the other bases of the interval = val (trackIO.py:198-201)
__global__ void paint_finalize_kernel(const int *__restrict__ owner, const int64_t *__restrict__ starts,
                                      const int32_t *__restrict__ vals, const int32_t *__restrict__ vals0,
                                      int64_t T, void *table, int K, int elem_bytes, int k)
{
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < T; x += (int64_t)gridDim.x * blockDim.x) {
        const int o = owner[x];
        if (o >= 0) store_elem(table, elem_bytes, x * K + k, x == starts[o] ? vals0[o] : vals[o]);
    }
}
__global__ void fill_column_kernel(void *table, int64_t T, int K, int elem_bytes, int k, int v)
{
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < T; x += (int64_t)gridDim.x * blockDim.x)
        store_elem(table, elem_bytes, x * K + k, v);
}

// ---------------------------------------------------------------- segmentation
// segmentTracks scans a region left to right keeping pi, the first column of the current segment
// ("first" comparison; "prev": the column to the left), and curLen, the columns since the last cut;
// column i starts a new segment when isNewSegment(pi, i, curLen) (segmentTracks.py:241-277).  The scan
// is serial but MEMORYLESS given the last cut, so it is chunked like the trellis: every chunk first
// scans as if a segment started at its first column; then chunks re-scan from the true last cut of
// their left neighbour until they produce a cut the speculative scan produced as well -- from there on
// the two scans coincide.  Passes repeat while some chunk's last cut changed (normally one repair pass).
struct SegParams {
    int K, elem_bytes, thresh, prev_mode;
    int64_t maxLen, fixLen;
    unsigned long long ignore_mask, cut_mask;      // K <= 64 tracks
};
#define SEG_CHUNK 256
#define SEG_MAXK 64

__device__ __forceinline__ bool seg_is_new(const SegParams &P, const void *table, int64_t pi, int64_t i, int64_t curLen)
{
    if (P.fixLen > 0) return curLen >= P.fixLen;
    if (P.maxLen > 0 && curLen >= P.maxLen) return true;
    int dif = 0;
    bool cut = false;
    for (int j = 0; j < P.K; ++j) {
        if ((P.ignore_mask >> j) & 1ull) continue;
        if (load_elem(table, P.elem_bytes, i * P.K + j) != load_elem(table, P.elem_bytes, pi * P.K + j)) {
            dif += 1;
            if ((P.cut_mask >> j) & 1ull) cut = true;
        }
    }
    return cut || dif > P.thresh;
}

// one thread per chunk.  region_of[c] / chunk bounds come from the chunk table (c0 = first row, c1 = end
// row, r0 = first row of the region).  pass 0: speculative (last cut = c0, except at a region start where
// that is the truth); pass > 0: from last_cut_in[c] (the left neighbour's last cut), until merged.
__global__ void segment_scan_kernel(SegParams P, const void *__restrict__ table, const int64_t *__restrict__ chunk_c0,
                                    const int64_t *__restrict__ chunk_c1, const int64_t *__restrict__ chunk_r0,
                                    int64_t nchunks, uint8_t *__restrict__ cut, const int64_t *__restrict__ last_in,
                                    int64_t *__restrict__ last_out, int pass, int *__restrict__ changed)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const int64_t c0 = chunk_c0[c], c1 = chunk_c1[c], r0 = chunk_r0[c];
    const bool region_start = c0 == r0;
    int64_t last;          // row of the last cut (the current segment's first column)
    int64_t i = c0;
    if (pass == 0) {
        last = c0;
        if (region_start) cut[c0] = 1;                 // a region always opens a segment (segmentTracks.py:212)
        else cut[c0] = 1;                              // speculation: a segment starts here
        i = c0 + 1;
    } else {
        if (region_start) { last_out[c] = last_in[c]; return; }      // exact in pass 0
        last = last_in[c - 1];                         // the truth about the left neighbour
        if (last == c0) { last_out[c] = last_in[c]; return; }        // cannot happen (c0 belongs to this chunk)
    }
    for (; i < c1; ++i) {
        const int64_t pi = P.prev_mode ? i - 1 : last;
        const bool isnew = seg_is_new(P, table, pi, i, i - last);
        if (pass > 0) {
            if (isnew && cut[i]) {                     // merged with the scan this chunk did before
                last_out[c] = last_in[c];
                return;
            }
            cut[i] = isnew ? 1 : 0;
        } else if (i > c0) {
            cut[i] = isnew ? 1 : 0;
        }
        if (isnew) last = i;
    }
    // reached the end of the chunk without merging (pass > 0) or simply done (pass 0)
    if (pass > 0 && last != last_in[c]) atomicExch(changed, 1);
    last_out[c] = last;
}
// (pass 0 writes cut[c0] = 1 for every chunk; in pass > 0 the re-scan starts AT c0, so a wrong
//  speculation there is overwritten like any other column)

// ---------------------------------------------------------------- compression
// one warp per segment: per track, the mode of the segment's values (the smallest value among the most
// frequent ones, as scipy.stats.mode, track.py:618-620), written to row `seg` of the output
__global__ void segment_mode_kernel(const void *__restrict__ table, int K, int elem_bytes,
                                    const int64_t *__restrict__ seg_off, int64_t nseg, int64_t T,
                                    unsigned long long mode_mask, void *__restrict__ out)
{
    extern __shared__ int hist_all[];             // per warp: 256 bins (uint8 tables)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int *hist = hist_all + warp * 256;
    for (int64_t s = (int64_t)blockIdx.x * wpb + warp; s < nseg; s += (int64_t)gridDim.x * wpb) {
        const int64_t a = seg_off[s], e = s + 1 < nseg ? seg_off[s + 1] : T;
        for (int k = 0; k < K; ++k) {
            int result;
            if (e - a == 1 || !((mode_mask >> k) & 1ull)) {
                result = load_elem(table, elem_bytes, a * K + k);
            } else {
                for (int b = lane; b < 256; b += 32) hist[b] = 0;
                __syncwarp();
                for (int64_t x = a + lane; x < e; x += 32) atomicAdd(hist + (load_elem(table, elem_bytes, x * K + k) & 255), 1);
                __syncwarp();
                int best = -1, arg = 0;
                for (int b = lane; b < 256; b += 32)
                    if (hist[b] > best) { best = hist[b]; arg = b; }     // ascending b per lane: first maximum kept
                for (int o = 16; o > 0; o >>= 1) {
                    const int ob = __shfl_xor_sync(TEHMM_FULL, best, o), oa = __shfl_xor_sync(TEHMM_FULL, arg, o);
                    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
                }
                result = arg;
                __syncwarp();
            }
            if (lane == 0) store_elem(out, elem_bytes, s * K + k, result);
        }
    }
}

struct IsZero {
    __device__ __forceinline__ int operator()(const uint8_t &m) const { return m == 0 ? 1 : 0; }
};

}   // namespace

extern "C" {

int tehmm_track_fill(tehmm_ctx *c, void *d_table, int64_t T, int K, int elem_bytes, int k, int32_t value)
{
    if (!c || !d_table || T < 0 || K <= 0 || k < 0 || k >= K) return terr(TEHMM_EINVAL, "bad argument");
    if (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4) return terr(TEHMM_EINVAL, "elem_bytes must be 1, 2 or 4");
    TCU(cudaSetDevice(tehmm_ctx_device(c)));
    if (T == 0) return TEHMM_OK;
    cudaStream_t st = (cudaStream_t)(uintptr_t)tehmm_ctx_stream(c);
    fill_column_kernel<<<(int)std::min<int64_t>((T + 255) / 256, 148 * 8), 256, 0, st>>>(d_table, T, K, elem_bytes, k, value);
    TCU(cudaGetLastError());
    return TEHMM_OK;
}

int tehmm_rasterize_intervals(tehmm_ctx *c, const int64_t *h_start, const int64_t *h_end, const int32_t *h_val,
                              const int32_t *h_val0, int64_t n, int64_t region_start, int64_t region_end,
                              void *d_table, int K, int elem_bytes, int k)
{
    if (!c || !d_table || n < 0 || (n > 0 && (!h_start || !h_end || !h_val || !h_val0))) return terr(TEHMM_EINVAL, "NULL argument");
    if (region_end <= region_start || K <= 0 || k < 0 || k >= K) return terr(TEHMM_EINVAL, "bad region or column");
    if (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4) return terr(TEHMM_EINVAL, "elem_bytes must be 1, 2 or 4");
    if (n > 0x7fffffff) return terr(TEHMM_ELIMIT, "more than 2^31 intervals in one call");
    TCU(cudaSetDevice(tehmm_ctx_device(c)));
    if (n == 0) return TEHMM_OK;
    cudaStream_t st = (cudaStream_t)(uintptr_t)tehmm_ctx_stream(c);
    const int64_t T = region_end - region_start;
    // clip to the region, in table coordinates; intervals outside keep their index but paint nothing
    std::vector<int64_t> s(n), e(n), piece_iv, piece_off;
    for (int64_t i = 0; i < n; ++i) {
        s[i] = std::max(h_start[i], region_start) - region_start;
        e[i] = std::min(h_end[i], region_end) - region_start;
        if (e[i] < s[i]) e[i] = s[i];
        if (e[i] - s[i] > 4 * PAINT_PIECE)
            for (int64_t o = 0; o < e[i] - s[i]; o += PAINT_PIECE) { piece_iv.push_back(i); piece_off.push_back(o); }
    }
    std::vector<int64_t> es(e);
    for (int64_t i = 0; i < n; ++i) if (e[i] - s[i] > 4 * PAINT_PIECE) es[i] = s[i];     // painted by the block kernel
    TmpBuf d_s, d_e, d_es, d_v, d_v0, d_owner, d_piv, d_poff;
    TCU(d_s.alloc((size_t)n * 8)); TCU(d_e.alloc((size_t)n * 8)); TCU(d_es.alloc((size_t)n * 8));
    TCU(d_v.alloc((size_t)n * 4)); TCU(d_v0.alloc((size_t)n * 4)); TCU(d_owner.alloc((size_t)T * 4));
    TCU(cudaMemcpyAsync(d_s.p, s.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    TCU(cudaMemcpyAsync(d_e.p, e.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    TCU(cudaMemcpyAsync(d_es.p, es.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    TCU(cudaMemcpyAsync(d_v.p, h_val, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    TCU(cudaMemcpyAsync(d_v0.p, h_val0, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    TCU(cudaMemsetAsync(d_owner.p, 0xff, (size_t)T * 4, st));
    paint_owner_kernel<<<(int)std::min<int64_t>((n * 32 + 255) / 256, 148 * 16), 256, 0, st>>>(d_s.as<int64_t>(), d_es.as<int64_t>(), n, d_owner.as<int>());
    TCU(cudaGetLastError());
    if (!piece_iv.empty()) {
        const int64_t np_ = (int64_t)piece_iv.size();
        TCU(d_piv.alloc((size_t)np_ * 8)); TCU(d_poff.alloc((size_t)np_ * 8));
        TCU(cudaMemcpyAsync(d_piv.p, piece_iv.data(), (size_t)np_ * 8, cudaMemcpyHostToDevice, st));
        TCU(cudaMemcpyAsync(d_poff.p, piece_off.data(), (size_t)np_ * 8, cudaMemcpyHostToDevice, st));
        paint_owner_long_kernel<<<(int)std::min<int64_t>(np_, 148 * 8), 256, 0, st>>>(d_s.as<int64_t>(), d_e.as<int64_t>(), d_piv.as<int64_t>(),
                                                                                       d_poff.as<int64_t>(), np_, d_owner.as<int>());
        TCU(cudaGetLastError());
    }
    paint_finalize_kernel<<<(int)std::min<int64_t>((T + 255) / 256, 148 * 8), 256, 0, st>>>(d_owner.as<int>(), d_s.as<int64_t>(), d_v.as<int32_t>(),
                                                                                         d_v0.as<int32_t>(), T, d_table, K, elem_bytes, k);
    TCU(cudaGetLastError());
    TCU(cudaStreamSynchronize(st));        // the temporaries go out of scope
    return TEHMM_OK;
}

int tehmm_segment_table(tehmm_ctx *c, const void *d_table, int64_t T, int K, int elem_bytes, int64_t nregions,
                        const int64_t *h_region_off, const uint8_t *h_ignore, const uint8_t *h_cut, int thresh,
                        int64_t maxLen, int64_t fixLen, int prev_mode, uint8_t *d_cut, int64_t *d_seg_off,
                        int64_t *h_nseg, int *h_passes)
{
    if (!c || !d_table || !h_region_off || !d_cut || !d_seg_off || !h_nseg) return terr(TEHMM_EINVAL, "NULL argument");
    if (T <= 0 || K <= 0 || K > SEG_MAXK || nregions <= 0) return terr(TEHMM_EINVAL, "bad shape (K <= 64)");
    if (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4) return terr(TEHMM_EINVAL, "elem_bytes must be 1, 2 or 4");
    if (h_region_off[0] != 0 || h_region_off[nregions] != T) return terr(TEHMM_EINVAL, "region offsets must span [0, T]");
    TCU(cudaSetDevice(tehmm_ctx_device(c)));
    cudaStream_t st = (cudaStream_t)(uintptr_t)tehmm_ctx_stream(c);
    SegParams P;
    P.K = K; P.elem_bytes = elem_bytes; P.thresh = thresh; P.prev_mode = prev_mode ? 1 : 0;
    P.maxLen = maxLen; P.fixLen = fixLen; P.ignore_mask = 0ull; P.cut_mask = 0ull;
    for (int j = 0; j < K; ++j) {
        if (h_ignore && h_ignore[j]) P.ignore_mask |= 1ull << j;
        if (h_cut && h_cut[j]) P.cut_mask |= 1ull << j;
    }
    std::vector<int64_t> c0, c1, r0;
    for (int64_t r = 0; r < nregions; ++r) {
        if (h_region_off[r + 1] <= h_region_off[r]) return terr(TEHMM_EINVAL, "empty region");
        for (int64_t a = h_region_off[r]; a < h_region_off[r + 1]; a += SEG_CHUNK) {
            c0.push_back(a); c1.push_back(std::min(h_region_off[r + 1], a + (int64_t)SEG_CHUNK)); r0.push_back(h_region_off[r]);
        }
    }
    const int64_t nch = (int64_t)c0.size();
    TmpBuf d_c0, d_c1, d_r0, d_la, d_lb, d_changed, d_tmp, d_n;
    TCU(d_c0.alloc((size_t)nch * 8)); TCU(d_c1.alloc((size_t)nch * 8)); TCU(d_r0.alloc((size_t)nch * 8));
    TCU(d_la.alloc((size_t)nch * 8)); TCU(d_lb.alloc((size_t)nch * 8)); TCU(d_changed.alloc(4)); TCU(d_n.alloc(8));
    TCU(cudaMemcpyAsync(d_c0.p, c0.data(), (size_t)nch * 8, cudaMemcpyHostToDevice, st));
    TCU(cudaMemcpyAsync(d_c1.p, c1.data(), (size_t)nch * 8, cudaMemcpyHostToDevice, st));
    TCU(cudaMemcpyAsync(d_r0.p, r0.data(), (size_t)nch * 8, cudaMemcpyHostToDevice, st));
    const int grid = (int)((nch + 127) / 128);
    int64_t *la = d_la.as<int64_t>(), *lb = d_lb.as<int64_t>();
    segment_scan_kernel<<<grid, 128, 0, st>>>(P, d_table, d_c0.as<int64_t>(), d_c1.as<int64_t>(), d_r0.as<int64_t>(), nch, d_cut, la, la, 0,
                                              d_changed.as<int>());
    TCU(cudaGetLastError());
    int passes = 1;
    for (;; ++passes) {
        TCU(cudaMemsetAsync(d_changed.p, 0, 4, st));
        segment_scan_kernel<<<grid, 128, 0, st>>>(P, d_table, d_c0.as<int64_t>(), d_c1.as<int64_t>(), d_r0.as<int64_t>(), nch, d_cut, la, lb,
                                                  passes, d_changed.as<int>());
        TCU(cudaGetLastError());
        int changed = 0;
        TCU(cudaMemcpyAsync(&changed, d_changed.p, 4, cudaMemcpyDeviceToHost, st));
        TCU(cudaStreamSynchronize(st));
        std::swap(la, lb);
        if (!changed) break;
        if (passes > nch + 1) return terr(TEHMM_ESTATE, "segmentation repair did not converge");
    }
    if (h_passes) *h_passes = passes + 1;
    // segment offsets = the rows whose flag is set
    size_t tmp_bytes = 0;
    cub::CountingInputIterator<int64_t> rows(0);
    cub::DeviceSelect::Flagged(nullptr, tmp_bytes, rows, d_cut, d_seg_off, d_n.as<int64_t>(), T, st);
    TCU(d_tmp.alloc(tmp_bytes));
    TCU(cub::DeviceSelect::Flagged(d_tmp.p, tmp_bytes, rows, d_cut, d_seg_off, d_n.as<int64_t>(), T, st));
    TCU(cudaMemcpyAsync(h_nseg, d_n.p, 8, cudaMemcpyDeviceToHost, st));
    TCU(cudaStreamSynchronize(st));
    return TEHMM_OK;
}

int tehmm_compress_segments(tehmm_ctx *c, const void *d_table, int64_t T, int K, int elem_bytes,
                            const int64_t *d_seg_off, int64_t nseg, const uint8_t *h_use_mode, void *d_out)
{
    if (!c || !d_table || !d_seg_off || !d_out) return terr(TEHMM_EINVAL, "NULL argument");
    if (T <= 0 || K <= 0 || K > SEG_MAXK || nseg <= 0) return terr(TEHMM_EINVAL, "bad shape (K <= 64)");
    if (elem_bytes != 1) {
        if (h_use_mode) for (int j = 0; j < K; ++j) if (h_use_mode[j]) return terr(TEHMM_ELIMIT, "the per-segment mode is implemented for uint8 tables");
        if (elem_bytes != 2 && elem_bytes != 4) return terr(TEHMM_EINVAL, "elem_bytes must be 1, 2 or 4");
    }
    TCU(cudaSetDevice(tehmm_ctx_device(c)));
    cudaStream_t st = (cudaStream_t)(uintptr_t)tehmm_ctx_stream(c);
    unsigned long long mask = 0ull;
    for (int j = 0; j < K; ++j) if (h_use_mode && h_use_mode[j]) mask |= 1ull << j;
    const int wpb = 8;
    const int64_t need = (nseg + wpb - 1) / wpb;
    segment_mode_kernel<<<(int)std::min<int64_t>(need, 148 * 8), wpb * 32, wpb * 256 * sizeof(int), st>>>(d_table, K, elem_bytes, d_seg_off, nseg, T, mask, d_out);
    TCU(cudaGetLastError());
    return TEHMM_OK;
}

int tehmm_run_sum(tehmm_ctx *c, const uint8_t *d_mask, int32_t *d_out, int64_t n)
{
    if (!c || !d_mask || !d_out || n < 0) return terr(TEHMM_EINVAL, "bad argument");
    TCU(cudaSetDevice(tehmm_ctx_device(c)));
    if (n == 0) return TEHMM_OK;
    if (n > 0x7fffffff) return terr(TEHMM_ELIMIT, "runSum is int32 (_track.pyx:8)");
    cudaStream_t st = (cudaStream_t)(uintptr_t)tehmm_ctx_stream(c);
    cub::TransformInputIterator<int32_t, IsZero, const uint8_t *> it(d_mask, IsZero());
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, d_out, (int)n, st);
    TmpBuf tmp;
    TCU(tmp.alloc(tmp_bytes));
    TCU(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, it, d_out, (int)n, st));
    TCU(cudaStreamSynchronize(st));
    return TEHMM_OK;
}

// ---------------------------------------------------------------- BED parsing (host)
// bedRead (trackIO.py:389-404) + what `intersectBed -a file -b interval | sortBed` leaves of it
// (trackIO.py:138-143): tab-separated lines with at least three columns, '#' lines skipped; with
// need_intersect only the parts of `chrom`'s intervals inside [start, end), stable-sorted by start.
// Values are returned as indices into the table of DISTINCT value strings in order of first
// appearance -- the order in which the reference's loop shows them to the track's value map.
struct tehmm_bed {
    std::vector<int64_t> s, e;
    std::vector<int32_t> vi;
    std::vector<std::string> uniq;
};

int tehmm_bed_open(const char *path, const char *chrom, int64_t start, int64_t end, int need_intersect, int sort,
                   int valcol, tehmm_bed **out)
{
    if (!path || !out || (need_intersect && !chrom)) return terr(TEHMM_EINVAL, "NULL argument");
    if (valcol != 0 && valcol != 3 && valcol != 4) return terr(TEHMM_EINVAL, "valcol must be 0 (none), 3 or 4");
    FILE *f = fopen(path, "rb");
    if (!f) return terr(TEHMM_EINVAL, "cannot open the BED file");
    std::string buf;
    {
        char tmp[1 << 16];
        size_t n;
        while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
    }
    fclose(f);
    struct Row { int64_t s, e; const char *v; int vlen; std::string chrom; };
    std::vector<Row> rows;
    const size_t clen = chrom ? strlen(chrom) : 0;
    const char *p = buf.data(), *endp = p + buf.size();
    while (p < endp) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(endp - p));
        const char *le = nl ? nl : endp;
        if (le > p && *p != '#') {
            const char *col[6];
            int len[6], nc = 0;
            const char *q = p;
            while (q <= le && nc < 6) {
                const char *t = (const char *)memchr(q, '\t', (size_t)(le - q));
                const char *ce = t ? t : le;
                col[nc] = q; len[nc] = (int)(ce - q); ++nc;
                if (!t) break;
                q = t + 1;
            }
            if (nc > 2) {
                Row r;
                r.s = strtoll(col[1], nullptr, 10);
                r.e = strtoll(col[2], nullptr, 10);
                r.v = nullptr; r.vlen = 0;
                if (valcol > 0) {
                    if (nc <= valcol || len[valcol] == 0) return terr(TEHMM_EINVAL, "BED line without the value column");
                    r.v = col[valcol]; r.vlen = len[valcol];
                }
                bool keep = true;
                if (need_intersect) {
                    keep = (size_t)len[0] == clen && memcmp(col[0], chrom, clen) == 0 && r.s < end && r.e > start;
                    if (keep) { r.s = std::max(r.s, start); r.e = std::min(r.e, end); }
                } else if (sort) {
                    r.chrom.assign(col[0], (size_t)len[0]);
                }
                if (keep) rows.push_back(std::move(r));
            }
        }
        if (!nl) break;
        p = nl + 1;
    }
    if (need_intersect) std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.s < b.s; });
    else if (sort) std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) { return a.chrom != b.chrom ? a.chrom < b.chrom : a.s < b.s; });
    tehmm_bed *h = new tehmm_bed();
    std::unordered_map<std::string, int32_t> idx;
    h->s.reserve(rows.size()); h->e.reserve(rows.size()); h->vi.reserve(rows.size());
    for (const Row &r : rows) {
        h->s.push_back(r.s); h->e.push_back(r.e);
        int32_t k = -1;
        if (valcol > 0) {
            std::string key(r.v, (size_t)r.vlen);
            auto it = idx.find(key);
            if (it == idx.end()) { k = (int32_t)h->uniq.size(); idx.emplace(key, k); h->uniq.push_back(key); }
            else k = it->second;
        }
        h->vi.push_back(k);
    }
    *out = h;
    return TEHMM_OK;
}
int64_t tehmm_bed_count(tehmm_bed *h) { return h ? (int64_t)h->s.size() : 0; }
int64_t tehmm_bed_nunique(tehmm_bed *h) { return h ? (int64_t)h->uniq.size() : 0; }
const char *tehmm_bed_unique(tehmm_bed *h, int64_t i) { return h && i >= 0 && i < (int64_t)h->uniq.size() ? h->uniq[(size_t)i].c_str() : nullptr; }
int tehmm_bed_fetch(tehmm_bed *h, int64_t *starts, int64_t *ends, int32_t *value_index)
{
    if (!h || !starts || !ends || !value_index) return terr(TEHMM_EINVAL, "NULL argument");
    const size_t n = h->s.size();
    if (n) {
        memcpy(starts, h->s.data(), n * 8); memcpy(ends, h->e.data(), n * 8); memcpy(value_index, h->vi.data(), n * 4);
    }
    return TEHMM_OK;
}
void tehmm_bed_close(tehmm_bed *h) { delete h; }

}   // extern "C"
