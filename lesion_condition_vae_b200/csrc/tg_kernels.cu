// tg_kernels.cu — kernels + C ABI (include/tractgeom.h) of the streamline-metrics path, sm_100a.
//
// Kernel 1  k_metrics_whole   one lane walks one polyline through the register pipeline of
//                             tg_device.cuh and writes its 17 metrics + keep flags.
// Kernel 2a k_bundle_tiles    per-tile partial moments of the 13 aggregated columns (tiles never
//                             straddle a bundle boundary) — ref:191-210 of tract_geom_proc.py.
// Kernel 2b k_bundle_final    one warp per bundle adds its tiles' partials in tile order, so the
//                             result does not depend on scheduling.
#include "tg_device.cuh"
#include "tg_grouped.cuh"
#include "tractgeom.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace tg {

// ------------------------------------------------------------------------------------------
// Kernel 1
// ------------------------------------------------------------------------------------------
constexpr int kMetricsThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kMetricsThreads)
k_metrics_whole(const T* __restrict__ xyz, const int64_t* __restrict__ offsets, const int64_t S,
                double* __restrict__ out, uint8_t* __restrict__ keep, const int* __restrict__ long_flag, const int64_t min_n) {
    // long_flag != nullptr: this launch only mops up the polylines k_metrics_grouped left (n > min_n), if any
    if (long_flag != nullptr && *long_flag == 0) return;
    const int64_t s = (int64_t)blockIdx.x * kMetricsThreads + threadIdx.x;
    if (s >= S) return;
    const int64_t o0 = __ldg(offsets + s), o1 = __ldg(offsets + s + 1);
    const int64_t n64 = o1 - o0;
    if (n64 <= min_n) return;
    if (n64 < 3) {                                           // ref:21  sl.shape[0] > 2
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
        for (int m = 0; m < TG_N_METRICS; ++m) out[(int64_t)m * S + s] = nan;
        keep[s] = 0;
        return;
    }
    const int n = (int)n64;
    const T* base = xyz + 3 * o0;
    double f0, f1, f2, g0, g1, g2, m0, m1, m2, e0, e1, e2;
    load_point(base, f0, f1, f2);
    load_point(base + 3, g0, g1, g2);
    load_point(base + 3 * (int64_t)(n >> 1), m0, m1, m2);
    load_point(base + 3 * (int64_t)(n - 1), e0, e1, e2);
    // reference direction r ~ first unit segment: any constant works (the dispersion is shift
    // invariant); this one makes the shifted sums small for nearly straight polylines.
    double rx = g0 - f0, ry = g1 - f1, rz = g2 - f2, rl, ri;
    norm_and_inv_eps(rx * rx + ry * ry + rz * rz, rl, ri);
    rx *= ri; ry *= ri; rz *= ri;
    if (!(finite_d(rx) && finite_d(ry) && finite_d(rz))) { rx = ry = rz = 0.0; }
    if (!(finite_d(m0) && finite_d(m1) && finite_d(m2))) { m0 = m1 = m2 = 0.0; }
    Acc A;
    acc_init(A);
    stream_chunk<T, true>(base, n, 0, n, rx, ry, rz, m0, m1, m2, A);
    keep[s] = (uint8_t)finalize_metrics(A, n, f0, f1, f2, e0, e1, e2, m0, m1, m2, out, S, s);
}

// float32 storage: exact upcast into a float64 scratch copy, then the float64 path (so a float32
// file gives bit-identical results to the same values stored as float64)
__global__ void __launch_bounds__(256)
k_upcast_f32(const float* __restrict__ src, double* __restrict__ dst, const int64_t count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < count; i += stride) {
        if (i + 3 < count && (((uintptr_t)(src + i)) & 15u) == 0) {
            const float4 v = *reinterpret_cast<const float4*>(src + i);
            dst[i] = (double)v.x; dst[i + 1] = (double)v.y; dst[i + 2] = (double)v.z; dst[i + 3] = (double)v.w;
        } else {
            for (int64_t j = i; j < count && j < i + 4; ++j) dst[j] = (double)src[j];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Kernel 2: bundle partial moments
// ------------------------------------------------------------------------------------------
constexpr int kBundleThreads = 256;
constexpr int kBundleTile = 4096;
constexpr int kNB = TG_N_BUNDLE_COLS;

struct TileDesc { int64_t begin, end; };

__constant__ int c_bundle_src[kNB];

__global__ void __launch_bounds__(kBundleThreads)
k_bundle_tiles(const double* __restrict__ out, const uint8_t* __restrict__ keep, const uint8_t* __restrict__ select,
               const int64_t S, const TileDesc* __restrict__ tiles, double* __restrict__ tsum, int64_t* __restrict__ tcnt) {
    const TileDesc t = tiles[blockIdx.x];
    double sum[kNB];
    int cnt[kNB];
    int kept = 0;
#pragma unroll
    for (int j = 0; j < kNB; ++j) { sum[j] = 0.0; cnt[j] = 0; }
    for (int64_t s = t.begin + threadIdx.x; s < t.end; s += kBundleThreads) {
        bool on = (keep[s] & TG_KEEP_BOTH) == TG_KEEP_BOTH;
        if (select != nullptr) on = on && (select[s] != 0);
        if (!on) continue;
        ++kept;
#pragma unroll
        for (int j = 0; j < kNB; ++j) {
            double v = __ldg(out + (int64_t)c_bundle_src[j] * S + s);
            if (v == v) { sum[j] += v; ++cnt[j]; }           // np.nanmean: skip NaN only, keep +-inf
        }
    }
    __shared__ double s_sum[kBundleThreads / 32][kNB];
    __shared__ int s_cnt[kBundleThreads / 32][kNB + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < kNB; ++j) {
        double v = sum[j];
        int c = cnt[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_down_sync(0xffffffffu, v, o);
            c += __shfl_down_sync(0xffffffffu, c, o);
        }
        if (lane == 0) { s_sum[warp][j] = v; s_cnt[warp][j + 1] = c; }
    }
    {
        int c = kept;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
        if (lane == 0) s_cnt[warp][0] = c;
    }
    __syncthreads();
    if (threadIdx.x < kNB) {
        double v = 0.0;
        for (int w = 0; w < kBundleThreads / 32; ++w) v += s_sum[w][threadIdx.x];
        tsum[(int64_t)blockIdx.x * kNB + threadIdx.x] = v;
    }
    if (threadIdx.x < kNB + 1) {
        int64_t c = 0;
        for (int w = 0; w < kBundleThreads / 32; ++w) c += s_cnt[w][threadIdx.x];
        tcnt[(int64_t)blockIdx.x * (kNB + 1) + threadIdx.x] = c;
    }
}

// one warp per bundle; tile_first[b]..tile_first[b+1] are its tiles
__global__ void __launch_bounds__(32)
k_bundle_final(const int64_t* __restrict__ tile_first, const double* __restrict__ tsum, const int64_t* __restrict__ tcnt,
               double* __restrict__ sums, int64_t* __restrict__ counts) {
    const int64_t b = blockIdx.x;
    const int64_t t0 = tile_first[b], t1 = tile_first[b + 1];
    const int lane = threadIdx.x;
    for (int j = 0; j < kNB + 1; ++j) {
        double v = 0.0;
        int64_t c = 0;
        for (int64_t t = t0 + lane; t < t1; t += 32) {
            if (j < kNB) v += tsum[t * kNB + j];
            c += tcnt[t * (kNB + 1) + j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_down_sync(0xffffffffu, v, o);
            c += __shfl_down_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            if (j < kNB) sums[b * kNB + j] = v;
            counts[b * (kNB + 1) + j] = c;
        }
    }
}

}  // namespace tg

// ==============================================================================================
// Host side: context, scratch, C ABI
// ==============================================================================================
static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof g_err, fmt, a, b);
    return code;
}
#define TG_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) return set_err(TG_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return TG_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_err(TG_E_NOMEM, "cudaMalloc(%s bytes) failed: %s", std::to_string(want).c_str(), cudaGetErrorString(e));
        }
        cap = want;
        return TG_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return TG_OK;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_err(TG_E_NOMEM, "cudaMallocHost(%s bytes) failed: %s", std::to_string(want).c_str(), cudaGetErrorString(e));
        }
        cap = want;
        return TG_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct tg_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t staged = nullptr;     // last H2D out of the pinned staging buffer
    bool staged_pending = false;
    int64_t launches = 0;
    int sm_count = 148;
    DevBuf d_qhead, d_hist, d_start, d_perm;   // length-binned queue scratch
    DevBuf d_xyz64;                            // float64 copy of float32 input
    bool grouped_ready = false;
    // bundle-reduce scratch
    PinBuf h_tiles;                   // TileDesc[nt] followed by int64 tile_first[B+1]
    DevBuf d_tiles, d_tsum, d_tcnt;
    // host-path scratch
    DevBuf d_xyz, d_off, d_out, d_keep, d_sums, d_counts;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int upload_bundle_src() {
    static thread_local int done_for = -1;
    int dev = -1;
    cudaGetDevice(&dev);
    if (done_for == dev) return TG_OK;
    TG_CUDA(cudaMemcpyToSymbol(tg::c_bundle_src, TG_BUNDLE_SOURCE, sizeof(int) * TG_N_BUNDLE_COLS));
    done_for = dev;
    return TG_OK;
}

}  // namespace

extern "C" {

int tg_abi_version(void) { return TG_ABI_VERSION; }
const char* tg_last_error(void) { return g_err; }

int tg_device_count(int* count) {
    if (!count) return set_err(TG_E_INVALID, "tg_device_count: null pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); *count = 0; return set_err(TG_E_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *count = n;
    return TG_OK;
}

int tg_create(int device, tg_context** out) {
    if (!out) return set_err(TG_E_INVALID, "tg_create: null ctx pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_err(TG_E_NODEVICE, "no CUDA device (%s); this library has no CPU path", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return set_err(TG_E_INVALID, "tg_create: device index out of range");
    cudaDeviceProp prop;
    TG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        return set_err(TG_E_NODEVICE, "device %s is not sm_100 (Blackwell B200); kernels are built for sm_100a only", prop.name);
    }
    DeviceGuard g(device);
    if (!g.ok) return set_err(TG_E_CUDA, "cudaSetDevice failed");
    tg_context* c = new (std::nothrow) tg_context();
    if (!c) return set_err(TG_E_NOMEM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    cudaError_t e1 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    cudaError_t e2 = cudaEventCreateWithFlags(&c->staged, cudaEventDisableTiming);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        delete c;
        return set_err(TG_E_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    }
    *out = c;
    return TG_OK;
}

int tg_destroy(tg_context* c) {
    if (!c) return TG_OK;
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    c->h_tiles.release();
    c->d_tiles.release(); c->d_tsum.release(); c->d_tcnt.release();
    c->d_xyz.release(); c->d_off.release(); c->d_out.release(); c->d_keep.release();
    c->d_sums.release(); c->d_counts.release();
    c->d_xyz64.release(); c->d_qhead.release(); c->d_hist.release(); c->d_start.release(); c->d_perm.release();
    cudaEventDestroy(c->staged);
    cudaStreamDestroy(c->stream);
    delete c;
    return TG_OK;
}

int tg_synchronize(tg_context* c) {
    if (!c) return set_err(TG_E_INVALID, "null context");
    DeviceGuard g(c->device);
    TG_CUDA(cudaStreamSynchronize(c->stream));
    return TG_OK;
}

int tg_stream(tg_context* c, void** stream) {
    if (!c || !stream) return set_err(TG_E_INVALID, "null argument");
    *stream = (void*)c->stream;
    return TG_OK;
}

int tg_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return set_err(TG_E_INVALID, "null pointer");
    *ptr = nullptr;
    if (bytes == 0) bytes = 1;
    cudaError_t e = cudaMallocHost(ptr, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(TG_E_NOMEM, "cudaMallocHost failed: %s", cudaGetErrorString(e)); }
    return TG_OK;
}
int tg_host_free(void* ptr) {
    if (!ptr) return TG_OK;
    TG_CUDA(cudaFreeHost(ptr));
    return TG_OK;
}

int tg_launch_count(tg_context* c, int64_t* launches) {
    if (!c || !launches) return set_err(TG_E_INVALID, "null argument");
    *launches = c->launches;
    return TG_OK;
}

int tg_metrics_csr_dev(tg_context* c, const void* d_xyz, int xyz_dtype, const int64_t* d_offsets, int64_t S, int64_t P,
                       double* d_out, uint8_t* d_keep, void* stream) {
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (S < 0 || P < 0) return set_err(TG_E_INVALID, "negative size");
    if (xyz_dtype != TG_F64 && xyz_dtype != TG_F32) return set_err(TG_E_INVALID, "xyz_dtype must be TG_F64 or TG_F32");
    if (S == 0) return TG_OK;
    if (!d_offsets || !d_out || !d_keep || (P > 0 && !d_xyz)) return set_err(TG_E_INVALID, "null device pointer");
    DeviceGuard g(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int64_t blocks = (S + tg::kMetricsThreads - 1) / tg::kMetricsThreads;
    if (blocks > 0x7fffffffLL) return set_err(TG_E_INVALID, "too many streamlines for one launch");
    if (xyz_dtype == TG_F32) {
        int rc0;
        if ((rc0 = c->d_xyz64.reserve(sizeof(double) * 3 * (size_t)P + 64))) return rc0;
        if (P > 0) {
            const int64_t count = 3 * P;
            const int64_t want = (count / 4 + 255) / 256;
            const unsigned g = (unsigned)(want < (int64_t)c->sm_count * 16 ? (want > 0 ? want : 1) : (int64_t)c->sm_count * 16);
            tg::k_upcast_f32<<<g, 256, 0, st>>>((const float*)d_xyz, (double*)c->d_xyz64.p, count);
            c->launches += 1;
        }
        d_xyz = c->d_xyz64.p;
    }
    {
        if (((uintptr_t)d_xyz & 7u) != 0) return set_err(TG_E_INVALID, "xyz must be 8-byte aligned");
        if (S > 0x7fffffffLL) return set_err(TG_E_INVALID, "too many streamlines for one launch");
        if (!c->grouped_ready) {
            TG_CUDA(cudaFuncSetAttribute(tg::k_metrics_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::kGroupedSmem));
            c->grouped_ready = true;
        }
        // queue scratch: [flag | total] , hist/cursor, start, perm
        const int64_t n_windows = (S + tg::kWindow - 1) / tg::kWindow;
        int rc;
        if ((rc = c->d_qhead.reserve(64))) return rc;
        if ((rc = c->d_hist.reserve(sizeof(unsigned) * tg::kBins * (size_t)n_windows))) return rc;
        if ((rc = c->d_start.reserve(sizeof(int64_t) * tg::kBins * (size_t)n_windows))) return rc;
        if ((rc = c->d_perm.reserve(sizeof(uint4) * (size_t)S))) return rc;
        int* d_flag = (int*)c->d_qhead.p;
        int64_t* d_total = (int64_t*)((char*)c->d_qhead.p + 8);
        unsigned* d_hist = (unsigned*)c->d_hist.p;
        int64_t* d_start = (int64_t*)c->d_start.p;
        uint4* d_perm = (uint4*)c->d_perm.p;
        TG_CUDA(cudaMemsetAsync(c->d_qhead.p, 0, 64, st));
        TG_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned) * tg::kBins * (size_t)n_windows, st));
        const unsigned seg_grid = (unsigned)((S + tg::kBinSeg - 1) / tg::kBinSeg);
        tg::k_bin_count<<<seg_grid, tg::kBinThreads, 0, st>>>(d_offsets, S, d_hist);
        tg::k_bin_scan<<<1, 1024, 0, st>>>(d_hist, n_windows, d_start, d_total);
        tg::k_bin_scatter<<<seg_grid, tg::kBinThreads, 0, st>>>(d_offsets, S, d_hist, d_start, d_perm, d_out, d_keep, d_flag);
        const int64_t groups = (S + 31) / 32;
        const int64_t ctas = (groups + tg::kWarpsPerCta - 1) / tg::kWarpsPerCta;
        const unsigned grid = (unsigned)(ctas < c->sm_count ? ctas : c->sm_count);
        tg::k_metrics_grouped<<<grid, tg::kGroupedThreads, tg::kGroupedSmem, st>>>((const double*)d_xyz, P, S, d_perm, d_total, (unsigned long long*)((char*)c->d_qhead.p + 16), d_out, d_keep);
        tg::k_metrics_whole<double><<<(unsigned)blocks, tg::kMetricsThreads, 0, st>>>((const double*)d_xyz, d_offsets, S, d_out, d_keep, d_flag, (int64_t)tg::kMaxGroupedN);
        c->launches += 5;
    }
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_bundle_reduce_dev(tg_context* c, const double* d_out, const uint8_t* d_keep, const uint8_t* d_select, int64_t S,
                         const int64_t* h_bo, int64_t B, double* d_sums, int64_t* d_counts, void* stream) {
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (S < 0 || B < 0) return set_err(TG_E_INVALID, "negative size");
    if (B == 0) return TG_OK;
    if (!h_bo || !d_sums || !d_counts || (S > 0 && (!d_out || !d_keep))) return set_err(TG_E_INVALID, "null pointer");
    if (h_bo[0] < 0 || h_bo[B] > S) return set_err(TG_E_INVALID, "bundle_offsets out of range");
    for (int64_t b = 0; b < B; ++b)
        if (h_bo[b + 1] < h_bo[b]) return set_err(TG_E_INVALID, "bundle_offsets must be non-decreasing");
    DeviceGuard g(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = upload_bundle_src();
    if (rc) return rc;

    int64_t nt = 0;
    for (int64_t b = 0; b < B; ++b) nt += (h_bo[b + 1] - h_bo[b] + tg::kBundleTile - 1) / tg::kBundleTile;
    const size_t tiles_bytes = sizeof(tg::TileDesc) * (size_t)nt;
    const size_t first_bytes = sizeof(int64_t) * (size_t)(B + 1);
    if (c->staged_pending) { TG_CUDA(cudaEventSynchronize(c->staged)); c->staged_pending = false; }
    if ((rc = c->h_tiles.reserve(tiles_bytes + first_bytes))) return rc;
    if ((rc = c->d_tiles.reserve(tiles_bytes + first_bytes))) return rc;
    if ((rc = c->d_tsum.reserve(sizeof(double) * tg::kNB * (size_t)(nt + 1)))) return rc;
    if ((rc = c->d_tcnt.reserve(sizeof(int64_t) * (tg::kNB + 1) * (size_t)(nt + 1)))) return rc;
    tg::TileDesc* ht = (tg::TileDesc*)c->h_tiles.p;
    int64_t* hf = (int64_t*)((char*)c->h_tiles.p + tiles_bytes);
    int64_t t = 0;
    for (int64_t b = 0; b < B; ++b) {
        hf[b] = t;
        for (int64_t s = h_bo[b]; s < h_bo[b + 1]; s += tg::kBundleTile) {
            ht[t].begin = s;
            ht[t].end = (s + tg::kBundleTile < h_bo[b + 1]) ? s + tg::kBundleTile : h_bo[b + 1];
            ++t;
        }
    }
    hf[B] = t;
    TG_CUDA(cudaMemcpyAsync(c->d_tiles.p, c->h_tiles.p, tiles_bytes + first_bytes, cudaMemcpyHostToDevice, st));
    TG_CUDA(cudaEventRecord(c->staged, st));
    c->staged_pending = true;
    const tg::TileDesc* dt = (const tg::TileDesc*)c->d_tiles.p;
    const int64_t* df = (const int64_t*)((const char*)c->d_tiles.p + tiles_bytes);
    if (nt > 0x7fffffffLL || B > 0x7fffffffLL) return set_err(TG_E_INVALID, "too many tiles/bundles for one launch");
    if (nt > 0) {
        tg::k_bundle_tiles<<<(unsigned)nt, tg::kBundleThreads, 0, st>>>(d_out, d_keep, d_select, S, dt, (double*)c->d_tsum.p, (int64_t*)c->d_tcnt.p);
        c->launches += 1;
    }
    tg::k_bundle_final<<<(unsigned)B, 32, 0, st>>>(df, (const double*)c->d_tsum.p, (const int64_t*)c->d_tcnt.p, d_sums, d_counts);
    c->launches += 1;
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_metrics_csr_host(tg_context* c, const void* h_xyz, int xyz_dtype, const int64_t* h_off, int64_t S, int64_t P,
                        const int64_t* h_bo, int64_t B, double* h_out, uint8_t* h_keep, double* h_sums, int64_t* h_counts) {
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (S < 0 || P < 0 || B < 0) return set_err(TG_E_INVALID, "negative size");
    if (xyz_dtype != TG_F64 && xyz_dtype != TG_F32) return set_err(TG_E_INVALID, "xyz_dtype must be TG_F64 or TG_F32");
    if (!h_off) return set_err(TG_E_INVALID, "null offsets");
    if (P > 0 && !h_xyz) return set_err(TG_E_INVALID, "null xyz");
    if (B > 0 && (!h_bo || !h_sums || !h_counts)) return set_err(TG_E_INVALID, "null bundle argument");
    if (h_off[0] < 0 || h_off[S] > P) return set_err(TG_E_INVALID, "offsets out of range of the point array");
    for (int64_t s = 0; s < S; ++s) {
        int64_t n = h_off[s + 1] - h_off[s];
        if (n < 0) return set_err(TG_E_INVALID, "offsets must be non-decreasing");
        if (n > 0x7ffffff0LL) return set_err(TG_E_INVALID, "a polyline has more than 2^31-16 points");
    }
    DeviceGuard g(c->device);
    const size_t esz = xyz_dtype == TG_F64 ? 8 : 4;
    int rc;
    if ((rc = c->d_xyz.reserve(esz * 3 * (size_t)P))) return rc;
    if ((rc = c->d_off.reserve(sizeof(int64_t) * (size_t)(S + 1)))) return rc;
    if ((rc = c->d_out.reserve(sizeof(double) * TG_N_METRICS * (size_t)S))) return rc;
    if ((rc = c->d_keep.reserve((size_t)S))) return rc;
    if ((rc = c->d_sums.reserve(sizeof(double) * tg::kNB * (size_t)B))) return rc;
    if ((rc = c->d_counts.reserve(sizeof(int64_t) * (tg::kNB + 1) * (size_t)B))) return rc;
    cudaStream_t st = c->stream;
    if (P > 0) TG_CUDA(cudaMemcpyAsync(c->d_xyz.p, h_xyz, esz * 3 * (size_t)P, cudaMemcpyHostToDevice, st));
    TG_CUDA(cudaMemcpyAsync(c->d_off.p, h_off, sizeof(int64_t) * (size_t)(S + 1), cudaMemcpyHostToDevice, st));
    if ((rc = tg_metrics_csr_dev(c, c->d_xyz.p, xyz_dtype, (const int64_t*)c->d_off.p, S, P, (double*)c->d_out.p, (uint8_t*)c->d_keep.p, st))) return rc;
    if (B > 0) {
        if ((rc = tg_bundle_reduce_dev(c, (const double*)c->d_out.p, (const uint8_t*)c->d_keep.p, nullptr, S, h_bo, B,
                                       (double*)c->d_sums.p, (int64_t*)c->d_counts.p, st))) return rc;
        TG_CUDA(cudaMemcpyAsync(h_sums, c->d_sums.p, sizeof(double) * tg::kNB * (size_t)B, cudaMemcpyDeviceToHost, st));
        TG_CUDA(cudaMemcpyAsync(h_counts, c->d_counts.p, sizeof(int64_t) * (tg::kNB + 1) * (size_t)B, cudaMemcpyDeviceToHost, st));
    }
    if (h_out && S > 0) TG_CUDA(cudaMemcpyAsync(h_out, c->d_out.p, sizeof(double) * TG_N_METRICS * (size_t)S, cudaMemcpyDeviceToHost, st));
    if (h_keep && S > 0) TG_CUDA(cudaMemcpyAsync(h_keep, c->d_keep.p, (size_t)S, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    return TG_OK;
}

}  // extern "C"
