// tg_kernels.cu — kernels + C ABI (include/tractgeom.h) of the streamline-metrics path, sm_100a.
//
// Kernel 1  (tg_grouped.cuh)  k_bin_count / k_bin_scan / k_bin_scatter: length-binned work queue;
//                             k_metrics_grouped: one polyline per lane, 17 metrics + keep flags.
//            k_metrics_long   polylines too long for the queue bins: one warp per polyline.
// Kernel 2a k_bundle_tiles    per-tile partial moments of the 13 aggregated columns (tiles never
//                             straddle a bundle boundary) — ref:191-210 of tract_geom_proc.py.
// Kernel 2b k_bundle_final    one CTA per bundle adds its tiles' partials in a fixed order, so the
//                             result does not depend on scheduling; optionally as ONE row of 27 doubles
//                             {13 sums | kept rows | 13 non-NaN counts}: the multi-GPU all-gather payload.
// Kernel 0  k_decode_points   point storage other than native float64 (float32; the big-endian float32 /
//                             float64 of a binary VTK file) -> float64, exactly, at HBM speed.
// Kernel 3  k_spread_*        opt-in np.nanstd / nanmin / nanmax of the bundle columns.
// Kernel 4  k_resample        arc-length resampling to K nodes (N4), a polyline staged by ONE TMA bulk copy.
// Host      tg_metrics_csr_host: chunked H2D || kernels || D2H; tg_batch_*: files pushed as they are parsed,
//           one device call per batch; tg_vtk_* / tg_parse_ascii_*: ingest helpers (no device).
#include "tg_device.cuh"
#include "tg_grouped.cuh"
#include "tractgeom.h"

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler (nsys, ncu --nvtx) is attached

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <algorithm>
#include <charconv>

namespace tg {

// ------------------------------------------------------------------------------------------
// Long polylines (more than kMaxGroupedN points): one WARP per polyline.  Every lane streams one
// contiguous chunk through the exact general pipeline of tg_device.cuh (with its 3-point halo),
// the 32 partial sums are merged pairwise in index order (Chan's formula for the curvature
// moments) and lane 0 finalises.  Rare by construction; keeps heavy-tailed tractograms balanced.
// ------------------------------------------------------------------------------------------
constexpr int kLongThreads = 128;

__device__ __forceinline__ double shfl_down_d(double v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
__device__ __forceinline__ void acc_shfl_down(const Acc& A, Acc& B, int o) {
    B.L = shfl_down_d(A.L, o); B.th = shfl_down_d(A.th, o);
    B.w0 = shfl_down_d(A.w0, o); B.w1 = shfl_down_d(A.w1, o); B.w2 = shfl_down_d(A.w2, o); B.ww = shfl_down_d(A.ww, o);
    B.q0 = shfl_down_d(A.q0, o); B.q1 = shfl_down_d(A.q1, o); B.q2 = shfl_down_d(A.q2, o);
    B.q00 = shfl_down_d(A.q00, o); B.q01 = shfl_down_d(A.q01, o); B.q02 = shfl_down_d(A.q02, o);
    B.q11 = shfl_down_d(A.q11, o); B.q12 = shfl_down_d(A.q12, o); B.q22 = shfl_down_d(A.q22, o);
    B.mn0 = shfl_down_d(A.mn0, o); B.mn1 = shfl_down_d(A.mn1, o); B.mn2 = shfl_down_d(A.mn2, o);
    B.mx0 = shfl_down_d(A.mx0, o); B.mx1 = shfl_down_d(A.mx1, o); B.mx2 = shfl_down_d(A.mx2, o);
    B.kK = shfl_down_d(A.kK, o); B.k1 = shfl_down_d(A.k1, o); B.k2 = shfl_down_d(A.k2, o);
    B.en = shfl_down_d(A.en, o); B.ta = shfl_down_d(A.ta, o);
    B.kn = __shfl_down_sync(0xffffffffu, A.kn, o); B.tn = __shfl_down_sync(0xffffffffu, A.tn, o);
    B.absmax_hi = __shfl_down_sync(0xffffffffu, A.absmax_hi, o);
}

__global__ void __launch_bounds__(kLongThreads)
k_metrics_long(const double* __restrict__ xyz, const int64_t* __restrict__ offsets, const int64_t ld,
               const unsigned* __restrict__ long_list_end, const int* __restrict__ n_long,
               double* __restrict__ out, uint8_t* __restrict__ keep) {
    const int count = *n_long;
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (kLongThreads / 32);
    for (int idx = blockIdx.x * (kLongThreads / 32) + (threadIdx.x >> 5); idx < count; idx += warps_total) {
        const int64_t s = (int64_t)long_list_end[-1 - (int64_t)idx];           // ids are stored downwards from the end of the queue buffer
        const int64_t o0 = __ldg(offsets + s), o1 = __ldg(offsets + s + 1);
        const int n = (int)min(o1 - o0, (int64_t)0x7ffffff0);
        const double* base = xyz + 3 * o0;
        double f0, f1, f2, g0, g1, g2, m0, m1, m2, e0, e1, e2;
        load_point(base, f0, f1, f2);
        load_point(base + 3, g0, g1, g2);
        load_point(base + 3 * (int64_t)(n >> 1), m0, m1, m2);
        load_point(base + 3 * (int64_t)(n - 1), e0, e1, e2);
        double rx = g0 - f0, ry = g1 - f1, rz = g2 - f2, rl, ri;
        norm_and_inv_eps(rx * rx + ry * ry + rz * rz, rl, ri);
        rx *= ri; ry *= ri; rz *= ri;
        if (!(finite_d(rx) && finite_d(ry) && finite_d(rz))) { rx = ry = rz = 0.0; }
        if (!(finite_d(m0) && finite_d(m1) && finite_d(m2))) { m0 = m1 = m2 = 0.0; }
        const int cs = (n + 31) / 32;
        const int c0 = min(lane * cs, n), c1 = min(c0 + cs, n);
        Acc A, B;
        acc_init(A);
        if (c0 < c1) stream_chunk<double, false>(base, n, c0, c1, rx, ry, rz, m0, m1, m2, A);
#pragma unroll 1
        for (int o = 1; o < 32; o <<= 1) {
            acc_shfl_down(A, B, o);
            if ((lane & (2 * o - 1)) == 0) acc_merge(A, B);
        }
        if (lane == 0) keep[s] = (uint8_t)finalize_metrics(A, n, f0, f1, f2, e0, e1, e2, m0, m1, m2, out, ld, s);
        __syncwarp();
    }
}

// Point storage other than native float64 -> a float64 scratch copy, then the float64 path: float32 is upcast exactly
// (a float32 file gives bit-identical results to the same values stored as float64), and the BIG-ENDIAN
// forms are what a legacy binary VTK file holds (`POINTS n float|double`): the loader ships the file's bytes as they
// are — into pinned memory, no per-point work on the host — and the byte swap happens here, at HBM speed.
__device__ __forceinline__ double decode_f32(uint32_t w, const bool swap) {
    if (swap) w = __byte_perm(w, 0u, 0x0123);
    return (double)__uint_as_float(w);
}
__device__ __forceinline__ double decode_f64(const uint2 w, const bool swap) {
    // little-endian memory order: w.x = bytes 0..3, w.y = bytes 4..7; big-endian value = reverse all 8 bytes
    const uint32_t lo = swap ? __byte_perm(w.y, 0u, 0x0123) : w.x;
    const uint32_t hi = swap ? __byte_perm(w.x, 0u, 0x0123) : w.y;
    return __hiloint2double((int)hi, (int)lo);
}
template <bool F32>
__global__ void __launch_bounds__(256)
k_decode_points(const void* __restrict__ src, double* __restrict__ dst, const int64_t count, const int swap) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < count; i += stride) {
        if (F32) {
            const uint32_t* s = (const uint32_t*)src + i;
            if (i + 3 < count && (((uintptr_t)s) & 15u) == 0) {
                const uint4 v = *reinterpret_cast<const uint4*>(s);
                dst[i] = decode_f32(v.x, swap); dst[i + 1] = decode_f32(v.y, swap); dst[i + 2] = decode_f32(v.z, swap); dst[i + 3] = decode_f32(v.w, swap);
            } else {
                for (int64_t j = i; j < count && j < i + 4; ++j) dst[j] = decode_f32(((const uint32_t*)src)[j], swap);
            }
        } else {
            const uint2* s = (const uint2*)src + i;      // 8-byte aligned by contract
            for (int64_t j = 0; j < 4 && i + j < count; ++j) dst[i + j] = decode_f64(s[j], swap);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Kernel 2: bundle partial moments
// ------------------------------------------------------------------------------------------
constexpr int kBundleThreads = 256;
constexpr int kBundleTile = 4096;
constexpr int kNB = TG_N_BUNDLE_COLS;

struct TileDesc { int64_t begin, end, bundle; };

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

// one CTA per bundle; tile_first[b]..tile_first[b+1] are its tiles.  Thread t adds tiles t, t+256, ... in
// order, then the 256 partials are added in a fixed tree: the result does not depend on scheduling.
constexpr int kFinalThreads = 256;
// `packed` != nullptr: the bundle's partial moments as ONE row of 27 doubles {13 sums | kept rows | 13 non-NaN counts}
// (counts < 2^53 are exact) — the payload of the multi-GPU all-gather, written by this kernel so that
// the collective can follow it on the stream with nothing in between.
__global__ void __launch_bounds__(kFinalThreads)
k_bundle_final(const int64_t* __restrict__ tile_first, const double* __restrict__ tsum, const int64_t* __restrict__ tcnt,
               double* __restrict__ sums, int64_t* __restrict__ counts, double* __restrict__ packed) {
    __shared__ double s_v[kFinalThreads];
    __shared__ long long s_c[kFinalThreads];
    const int64_t b = blockIdx.x;
    const int64_t t0 = tile_first[b], t1 = tile_first[b + 1];
    for (int j = 0; j < kNB + 1; ++j) {
        double v = 0.0;
        long long c = 0;
        for (int64_t t = t0 + threadIdx.x; t < t1; t += kFinalThreads) {
            if (j < kNB) v += tsum[t * kNB + j];
            c += tcnt[t * (kNB + 1) + j];
        }
        s_v[threadIdx.x] = v; s_c[threadIdx.x] = c;
        __syncthreads();
        for (int o = kFinalThreads / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) { s_v[threadIdx.x] += s_v[threadIdx.x + o]; s_c[threadIdx.x] += s_c[threadIdx.x + o]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (packed != nullptr) {
                if (j < kNB) packed[b * (2 * kNB + 1) + j] = s_v[0];
                packed[b * (2 * kNB + 1) + kNB + j] = (double)s_c[0];          // [13] = kept rows, [14 + c] = non-NaN entries of column c
            } else {
                if (j < kNB) sums[b * kNB + j] = s_v[0];
                counts[b * (kNB + 1) + j] = s_c[0];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Kernel 3 (opt-in): spread of the 13 bundle columns — population standard
// deviation (np.nanstd, what ref:193 `_safe_std` defines and never calls), minimum, maximum.
// Two-pass like numpy: deviations from the bundle mean that kernel 2 produced.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBundleThreads)
k_spread_tiles(const double* __restrict__ out, const uint8_t* __restrict__ keep, const uint8_t* __restrict__ select,
               const int64_t S, const TileDesc* __restrict__ tiles, const double* __restrict__ sums, const int64_t* __restrict__ counts,
               double* __restrict__ tspread) {
    const TileDesc t = tiles[blockIdx.x];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double mean[kNB], ssq[kNB], lo[kNB], hi[kNB];
#pragma unroll
    for (int j = 0; j < kNB; ++j) {
        const int64_t c = counts[t.bundle * (kNB + 1) + 1 + j];
        mean[j] = c > 0 ? sums[t.bundle * kNB + j] / (double)c : 0.0;
        ssq[j] = 0.0; lo[j] = inf; hi[j] = -inf;
    }
    for (int64_t s = t.begin + threadIdx.x; s < t.end; s += kBundleThreads) {
        bool on = (keep[s] & TG_KEEP_BOTH) == TG_KEEP_BOTH;
        if (select != nullptr) on = on && (select[s] != 0);
        if (!on) continue;
#pragma unroll
        for (int j = 0; j < kNB; ++j) {
            const double v = __ldg(out + (int64_t)c_bundle_src[j] * S + s);
            if (v == v) {
                const double d = v - mean[j];                  // inf - inf = NaN, as in numpy
                ssq[j] += d * d;
                lo[j] = fmin(lo[j], v); hi[j] = fmax(hi[j], v);
            }
        }
    }
    __shared__ double s_part[kBundleThreads / 32][kNB][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < kNB; ++j) {
        double v = ssq[j], a = lo[j], b = hi[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_down_sync(0xffffffffu, v, o);
            a = fmin(a, __shfl_down_sync(0xffffffffu, a, o));
            b = fmax(b, __shfl_down_sync(0xffffffffu, b, o));
        }
        if (lane == 0) { s_part[warp][j][0] = v; s_part[warp][j][1] = a; s_part[warp][j][2] = b; }
    }
    __syncthreads();
    if (threadIdx.x < kNB) {
        double v = 0.0, a = inf, b = -inf;
        for (int w = 0; w < kBundleThreads / 32; ++w) {
            v += s_part[w][threadIdx.x][0];
            a = fmin(a, s_part[w][threadIdx.x][1]);
            b = fmax(b, s_part[w][threadIdx.x][2]);
        }
        double* dst = tspread + ((int64_t)blockIdx.x * kNB + threadIdx.x) * 3;
        dst[0] = v; dst[1] = a; dst[2] = b;
    }
}

// one CTA per bundle, fixed-order tree like k_bundle_final; spread[b][j] = {std, min, max}, NaN when the column
// has no non-NaN entry.  A NaN partial (inf - inf) must survive the tree: fmin/fmax would drop it, the sum keeps it.
__global__ void __launch_bounds__(kFinalThreads)
k_spread_final(const int64_t* __restrict__ tile_first, const double* __restrict__ tspread, const int64_t* __restrict__ counts,
               double* __restrict__ spread) {
    __shared__ double s_v[kFinalThreads], s_a[kFinalThreads], s_b[kFinalThreads];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const int64_t b = blockIdx.x;
    const int64_t t0 = tile_first[b], t1 = tile_first[b + 1];
    for (int j = 0; j < kNB; ++j) {
        double v = 0.0, lo = inf, hi = -inf;
        for (int64_t t = t0 + threadIdx.x; t < t1; t += kFinalThreads) {
            const double* src = tspread + (t * kNB + j) * 3;
            v += src[0]; lo = fmin(lo, src[1]); hi = fmax(hi, src[2]);
        }
        s_v[threadIdx.x] = v; s_a[threadIdx.x] = lo; s_b[threadIdx.x] = hi;
        __syncthreads();
        for (int o = kFinalThreads / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                s_v[threadIdx.x] += s_v[threadIdx.x + o];
                s_a[threadIdx.x] = fmin(s_a[threadIdx.x], s_a[threadIdx.x + o]);
                s_b[threadIdx.x] = fmax(s_b[threadIdx.x], s_b[threadIdx.x + o]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int64_t c = counts[b * (kNB + 1) + 1 + j];
            double* dst = spread + (b * kNB + j) * 3;
            dst[0] = c > 0 ? sqrt(s_v[0] / (double)c) : nan;
            dst[1] = c > 0 ? s_a[0] : nan;
            dst[2] = c > 0 ? s_b[0] : nan;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Kernel 4: arc-length resampling of every polyline to K nodes — the ragged-to-fixed
// step that would feed /root/reference/src/vae/data_loader.py:94-100 (exactly 100 `point_id`s per
// streamline; the reference ships no producer).  Node k lies at arc length k L/(K-1) on the polyline,
// linearly interpolated inside its segment; node K-1 is the last point itself.
// One WARP per polyline.  Staged path (k_resample, up to 129 points): the points wait in shared memory, pass 1
// is segment-parallel (cumulative lengths by warp scans), pass 2 node-parallel (binary search per node).
// Generic path (resample_generic: longer polylines, or ones at the very edge of the point array): points read
// from global memory twice, every lane owns one segment [c0, c1) per chunk and emits the nodes inside it;
// c1 of lane i and c0 of lane i+1 are the SAME double (shuffled, never recomputed), so every node is emitted
// exactly once.  Algorithmic traffic: 24 n bytes read, 24 K bytes written per polyline.
// ------------------------------------------------------------------------------------------
constexpr int kRsWarps = 8;
constexpr int kResampleThreads = kRsWarps * 32;
constexpr int kRsMaxN = 129;                              // staged path: polylines of up to 129 points (4 chunks of 32 segments)
constexpr int kRsInBytes = ((kRsMaxN * 24 + 8 + 15) / 16) * 16 + 16;   // 8 bytes of skew in front, whole 16-byte pieces
constexpr int kRsCumBytes = ((kRsMaxN * 8 + 15) / 16) * 16;            // cumulative length at every point
#ifndef TG_RESAMPLE_BULK
#define TG_RESAMPLE_BULK 1    // a polyline is contiguous and belongs to ONE warp: lane 0 stages it with ONE TMA bulk copy (cp.async.bulk,
#endif                        // SASS UBLKCP) completing on an mbarrier; 0 = the cp.async form (32 lanes x 16-byte pieces): 13.65 ms vs 13.05 ms per 10M
constexpr int kRsWarpSmem = 2 * kRsInBytes + kRsCumBytes + (TG_RESAMPLE_BULK ? 16 : 0);
constexpr int kResampleSmem = kRsWarps * kRsWarpSmem;
constexpr int kResampleCtasPerSm = 3;
static_assert(kRsInBytes % 16 == 0 && kRsCumBytes % 16 == 0, "16-byte aligned staging buffers");

__device__ __forceinline__ double warp_scan_inclusive(double v, const int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy) : "memory");
}
// no usable length: an empty polyline or a non-finite length -> NaN nodes; zero length (or a single point) -> the first point
__device__ __forceinline__ void resample_degenerate(const double* __restrict__ p, const int64_t n, const double L, const int K,
                                                    double* __restrict__ dst, const int lane) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const bool first = n > 0 && L == 0.0;
    const double x = first ? p[0] : nan, y = first ? p[1] : nan, z = first ? p[2] : nan;
    for (int k = lane; k < K; k += 32) { dst[3 * k] = x; dst[3 * k + 1] = y; dst[3 * k + 2] = z; }
}

// Nodes of the segment [c0, c1) that starts at a with direction d: node k lies at arc length k * step.
// `out` is the polyline's node table (shared or global memory).
__device__ __forceinline__ void resample_emit(const double c0, const double c1, const double step, const double inv_step, const int K,
                                              const double ax, const double ay, const double az,
                                              const double dx, const double dy, const double dz, double* __restrict__ out) {
    // first node index with k * step >= c0 (the product is only a guess: settle it with exact comparisons)
    int k = min(max(__double2int_ru(c0 * inv_step), 0), K);
    while (k > 0 && (double)(k - 1) * step >= c0) --k;
    while (k < K - 1 && (double)k * step < c0) ++k;
    const double inv = rcp_fast(c1 - c0);
    for (; k < K - 1; ++k) {
        const double t = (double)k * step;
        if (!(t < c1)) break;
        const double r = fmin(fmax((t - c0) * inv, 0.0), 1.0);
        out[3 * k] = ax + r * dx; out[3 * k + 1] = ay + r * dy; out[3 * k + 2] = az + r * dz;
    }
}

// any length, any K: points read from global memory twice (the second pass hits L1/L2), nodes stored directly
__device__ __noinline__ void resample_generic(const double* __restrict__ p, const int64_t n, const int K, double* __restrict__ dst, const int lane) {
    double part = 0.0;
    for (int64_t i = lane; i < n - 1; i += 32) {
        const double dx = p[3 * i + 3] - p[3 * i], dy = p[3 * i + 4] - p[3 * i + 1], dz = p[3 * i + 5] - p[3 * i + 2];
        part += sqrt_fast((dx * dx + dy * dy) + dz * dz);
    }
    const double L = warp_sum(part);
    if (!(L > 0.0) || !(L < __longlong_as_double(0x7ff0000000000000LL))) { resample_degenerate(p, n, L, K, dst, lane); return; }
    const double step = L / (double)(K - 1);
    const double inv_step = (double)(K - 1) / L;
    double carry = 0.0;
    for (int64_t base = 0; base < n - 1; base += 32) {
        const int64_t i = base + lane;
        double seg = 0.0, ax = 0.0, ay = 0.0, az = 0.0, dx = 0.0, dy = 0.0, dz = 0.0;
        if (i < n - 1) {
            ax = p[3 * i]; ay = p[3 * i + 1]; az = p[3 * i + 2];
            dx = p[3 * i + 3] - ax; dy = p[3 * i + 4] - ay; dz = p[3 * i + 5] - az;
            seg = sqrt_fast((dx * dx + dy * dy) + dz * dz);
        }
        const double c1 = carry + warp_scan_inclusive(seg, lane);
        double c0 = __shfl_up_sync(0xffffffffu, c1, 1);
        if (lane == 0) c0 = carry;
        carry = __shfl_sync(0xffffffffu, c1, 31);
        if (i < n - 1 && c1 > c0) resample_emit(c0, c1, step, inv_step, K, ax, ay, az, dx, dy, dz, dst);
    }
    if (lane == 0) {                                               // the last node is the last point itself
        dst[3 * (K - 1)] = p[3 * (n - 1)]; dst[3 * (K - 1) + 1] = p[3 * (n - 1) + 1]; dst[3 * (K - 1) + 2] = p[3 * (n - 1) + 2];
    }
}

// Persistent grid, every warp walks polylines s, s + W, s + 2W, ...  The points of polyline s + W are in flight
// (one TMA bulk copy per polyline, issued by lane 0, completing on the buffer's mbarrier; 2 buffers per warp) while
// polyline s is resampled out of shared memory.
//   pass 1 (segment-parallel): lane l owns segment 32 c + l of chunk c; a warp scan per chunk leaves the
//           cumulative length at every point in shared memory;
//   pass 2 (node-parallel):    lane l owns node 32 r + l of round r: a branch-free binary search finds the last
//           point whose cumulative length is <= the node's arc length (the segment then has positive length),
//           the node is interpolated and stored — 3 x 8-byte stores per lane, 768 contiguous bytes per warp.
// No lane-divergent loops, and every node has exactly one owner by construction.
// [xyz_lo, xyz_hi) = bytes of the point array: a polyline whose 16-byte pieces would cross them takes the
// generic path, like the ones too long for the buffers.
__global__ void __launch_bounds__(kResampleThreads, kResampleCtasPerSm)
k_resample(const double* __restrict__ xyz, const uint64_t xyz_lo, const uint64_t xyz_hi, const int64_t* __restrict__ offsets,
           const int64_t S, const int K, double* __restrict__ nodes) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wsm = smem + warp * kRsWarpSmem;
    double* cum = (double*)(wsm + 2 * kRsInBytes);
    const uint32_t in_u32 = smem_u32(wsm);
    const uint64_t l2_stream = policy_point_reads();
    const int64_t W = (int64_t)gridDim.x * kRsWarps;
    const double inv_km1 = 1.0 / (double)(K - 1);
#if TG_RESAMPLE_BULK
    const uint32_t bar_u32 = in_u32 + 2 * kRsInBytes + kRsCumBytes;          // one mbarrier per input buffer, a single arrival each
    if (lane == 0) { tg::mbar_init(bar_u32, 1); tg::mbar_init(bar_u32 + 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t bar_phase = 0u;
#endif

    // stage polyline (o, n) into buffer b; returns the byte skew of its first point inside the buffer, or -1 if not staged
    auto stage = [&](const int64_t o, const int64_t n, const int b) -> int {
        if (n < 2 || n > kRsMaxN) return -1;
        const uint64_t base = (uint64_t)(uintptr_t)(xyz + 3 * o);
        const uint64_t a0 = base & ~(uint64_t)15;
        const int skew = (int)(base - a0);
        const int bytes = (skew + 24 * (int)n + 15) & ~15;
        if (a0 < xyz_lo || a0 + (uint64_t)bytes > xyz_hi) return -1;
        const uint32_t dst = in_u32 + b * kRsInBytes;
#if TG_RESAMPLE_BULK
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the buffer was last read through the generic proxy
            tg::mbar_arrive_tx(bar_u32 + 8 * b, (uint32_t)bytes);
            tg::bulk_g2s(dst, (const void*)(uintptr_t)a0, (uint32_t)bytes, bar_u32 + 8 * b, l2_stream);
        }
#else
        for (int j = lane * 16; j < bytes; j += 512) cp_async16(dst + j, (const unsigned char*)(uintptr_t)a0 + j, l2_stream);
#endif
        return skew;
    };

    int64_t s = (int64_t)blockIdx.x * kRsWarps + warp;
    int64_t o_cur = 0, n_cur = 0, o_nxt = 0, n_nxt = 0;
    if (s < S) { o_cur = __ldg(offsets + s); n_cur = __ldg(offsets + s + 1) - o_cur; }
    if (s + W < S) { o_nxt = __ldg(offsets + s + W); n_nxt = __ldg(offsets + s + W + 1) - o_nxt; }
    int skew_cur = (s < S) ? stage(o_cur, n_cur, 0) : -1;
    cp_async_commit();
    int buf = 0;
    for (; s < S; s += W) {
        // offsets of the polyline after next: in flight during this iteration
        int64_t o_nn = 0, n_nn = 0;
        if (s + 2 * W < S) { o_nn = __ldg(offsets + s + 2 * W); n_nn = __ldg(offsets + s + 2 * W + 1) - o_nn; }
        const int skew_nxt = (s + W < S) ? stage(o_nxt, n_nxt, buf ^ 1) : -1;
#if TG_RESAMPLE_BULK
        if (skew_cur >= 0) {                                               // staged: wait for the bulk copy into `buf`
            tg::mbar_wait(bar_u32 + 8 * buf, (bar_phase >> buf) & 1u);
            bar_phase ^= 1u << buf;
        }
#else
        cp_async_commit();
        cp_async_wait<1>();
#endif
        __syncwarp();
        const int n = (int)n_cur;
        double* dst = nodes + s * K * 3;
        if (skew_cur < 0) {
            const double* p = xyz + 3 * o_cur;
            if (n_cur <= 0) resample_degenerate(p, n_cur, 0.0, K, dst, lane);
            else resample_generic(p, n_cur, K, dst, lane);
        } else {
            const double* pin = (const double*)(wsm + buf * kRsInBytes + skew_cur);
            // ---- pass 1: cumulative length at every point
            double carry = 0.0;
            if (lane == 0) cum[0] = 0.0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (32 * c < n - 1) {                                      // warp-uniform
                    const int i = 32 * c + lane;
                    double seg = 0.0;
                    if (i < n - 1) {
                        const double dx = pin[3 * i + 3] - pin[3 * i], dy = pin[3 * i + 4] - pin[3 * i + 1], dz = pin[3 * i + 5] - pin[3 * i + 2];
                        seg = sqrt_fast((dx * dx + dy * dy) + dz * dz);
                    }
                    const double c1 = carry + warp_scan_inclusive(seg, lane);
                    if (i < n - 1) cum[i + 1] = c1;
                    carry = __shfl_sync(0xffffffffu, c1, 31);
                }
            }
            const double L = carry;
            __syncwarp();
            if (!(L > 0.0) || !(L < __longlong_as_double(0x7ff0000000000000LL))) {
                resample_degenerate(pin, n, L, K, dst, lane);
            } else {
                const double step = L * inv_km1;
                const int top = n - 2;                                     // last segment
                // ---- pass 2: one node per lane and round
                for (int k = lane; k < K; k += 32) {
                    const double t = (double)k * step;
                    int i = 0;                                             // last point with cum <= t, among 0 .. n-2
#pragma unroll
                    for (int h = 64; h > 0; h >>= 1) {
                        const int j = i + h;
                        const double cj = cum[min(j, top)];
                        i = (j <= top && cj <= t) ? j : i;
                    }
                    const double c0 = cum[i], c1 = cum[i + 1];
                    const double ax = pin[3 * i], ay = pin[3 * i + 1], az = pin[3 * i + 2];
                    const double bx = pin[3 * i + 3], by = pin[3 * i + 4], bz = pin[3 * i + 5];
                    double r = fmin(fmax((t - c0) * rcp_fast(c1 - c0), 0.0), 1.0);
                    const bool end = k == K - 1;                           // the last node is the last point itself
                    const double ex = pin[3 * (n - 1)], ey = pin[3 * (n - 1) + 1], ez = pin[3 * (n - 1) + 2];
                    const double x = end ? ex : ax + r * (bx - ax), y = end ? ey : ay + r * (by - ay), z = end ? ez : az + r * (bz - az);
                    dst[3 * k] = x; dst[3 * k + 1] = y; dst[3 * k + 2] = z;
                }
            }
        }
        __syncwarp();                                                      // buffer `buf` and cum are free again
        o_cur = o_nxt; n_cur = n_nxt; skew_cur = skew_nxt;
        o_nxt = o_nn; n_nxt = n_nn;
        buf ^= 1;
    }
    cp_async_wait<0>();
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

// NVTX range over a C-ABI call (the reference has no tracing; nsys/ncu timelines show these names)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

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
    cudaStream_t s_copy = nullptr, s_back = nullptr;   // H2D / D2H streams of the chunked host path
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr};
    bool grouped_ready = false, resample_ready = false;
    // bundle-reduce scratch
    PinBuf h_tiles;                   // TileDesc[nt] followed by int64 tile_first[B+1]
    DevBuf d_tiles, d_tsum, d_tcnt, d_tspread;
    // host-path scratch
    DevBuf d_xyz, d_off, d_out, d_keep, d_sums, d_counts, d_spread, d_nodes;
    // batch (tg_batch_*): raw bytes as pushed, their float64 decode, offsets assembled on the host
    DevBuf d_braw, d_bxyz;
    PinBuf h_boff;
    cudaEvent_t ev_push = nullptr;
    int64_t b_P = 0, b_S = 0, b_Pcap = -1, b_Scap = -1;
    size_t b_raw = 0;
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

// Tile table of a bundle partition: TileDesc[nt] (tiles never straddle bundles) followed by tile_first[B+1],
// built in pinned memory and copied to the device on `st`.
int plan_bundle_tiles(tg_context* c, const int64_t* h_bo, int64_t B, cudaStream_t st, int64_t* n_tiles,
                      const tg::TileDesc** d_tiles, const int64_t** d_first) {
    int64_t nt = 0;
    for (int64_t b = 0; b < B; ++b) nt += (h_bo[b + 1] - h_bo[b] + tg::kBundleTile - 1) / tg::kBundleTile;
    if (nt > 0x7fffffffLL || B > 0x7fffffffLL) return set_err(TG_E_INVALID, "too many tiles/bundles for one launch");
    const size_t tiles_bytes = sizeof(tg::TileDesc) * (size_t)nt;
    const size_t first_bytes = sizeof(int64_t) * (size_t)(B + 1);
    int rc;
    if (c->staged_pending) { TG_CUDA(cudaEventSynchronize(c->staged)); c->staged_pending = false; }
    if ((rc = c->h_tiles.reserve(tiles_bytes + first_bytes))) return rc;
    if ((rc = c->d_tiles.reserve(tiles_bytes + first_bytes))) return rc;
    tg::TileDesc* ht = (tg::TileDesc*)c->h_tiles.p;
    int64_t* hf = (int64_t*)((char*)c->h_tiles.p + tiles_bytes);
    int64_t t = 0;
    for (int64_t b = 0; b < B; ++b) {
        hf[b] = t;
        for (int64_t s = h_bo[b]; s < h_bo[b + 1]; s += tg::kBundleTile) {
            ht[t].begin = s;
            ht[t].end = (s + tg::kBundleTile < h_bo[b + 1]) ? s + tg::kBundleTile : h_bo[b + 1];
            ht[t].bundle = b;
            ++t;
        }
    }
    hf[B] = t;
    TG_CUDA(cudaMemcpyAsync(c->d_tiles.p, c->h_tiles.p, tiles_bytes + first_bytes, cudaMemcpyHostToDevice, st));
    TG_CUDA(cudaEventRecord(c->staged, st));
    c->staged_pending = true;
    *n_tiles = nt;
    *d_tiles = (const tg::TileDesc*)c->d_tiles.p;
    *d_first = (const int64_t*)((const char*)c->d_tiles.p + tiles_bytes);
    return TG_OK;
}

}  // namespace

extern "C" {

int tg_abi_version(void) { return TG_ABI_VERSION; }
#ifndef TG_BUILD_ID
#define TG_BUILD_ID "unidentified"
#endif
// "@(#)TG_BUILD_ID=" is the marker build.py looks for in the file's bytes
static const char k_build_id[] = "@(#)TG_BUILD_ID=" TG_BUILD_ID;
const char* tg_build_id(void) { return k_build_id + 16; }
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
    c->d_sums.release(); c->d_counts.release(); c->d_spread.release(); c->d_tspread.release(); c->d_nodes.release();
    c->d_xyz64.release(); c->d_qhead.release(); c->d_hist.release(); c->d_start.release(); c->d_perm.release();
    c->d_braw.release(); c->d_bxyz.release(); c->h_boff.release();
    if (c->ev_push) cudaEventDestroy(c->ev_push);
    if (c->s_copy) { cudaStreamDestroy(c->s_copy); cudaStreamDestroy(c->s_back); for (int i = 0; i < 2; ++i) { cudaEventDestroy(c->ev_h2d[i]); cudaEventDestroy(c->ev_comp[i]); } }
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

// Queue + streaming + long-polyline kernels for S polylines whose points are float64.
//   xyz      base such that polyline s starts at xyz + 3*offsets[s] (may be virtual: chunked host path)
//   lo, hi   byte range of the point array that may be read
//   ld       column stride of d_out (>= S); d_out / d_keep already point at this range's first polyline
static int launch_metrics_f64(tg_context* c, const double* xyz, uint64_t lo, uint64_t hi, const int64_t* d_offsets, int64_t S,
                              double* d_out, int64_t ld, uint8_t* d_keep, cudaStream_t st) {
    if (S > 0x7fffffffLL) return set_err(TG_E_INVALID, "too many streamlines for one launch");
    if (!c->grouped_ready) {
        TG_CUDA(cudaFuncSetAttribute(tg::k_metrics_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::kGroupedSmem));
        c->grouped_ready = true;
    }
    // queue scratch: [n_long | total | ticket | long_grouped], hist/cursor, start, queue records (+ long ids at the end)
    const int64_t n_windows = (S + tg::kWindow - 1) / tg::kWindow + 1;   // queue rows: row 0 (long polylines) + one per window
    int rc;
    if (!c->d_qhead.p) {
        if ((rc = c->d_qhead.reserve(64))) return rc;
        TG_CUDA(cudaMemsetAsync(c->d_qhead.p, 0, 64, st));
    }
    if ((rc = c->d_hist.reserve(sizeof(unsigned) * tg::kBins * (size_t)n_windows))) return rc;
    if ((rc = c->d_start.reserve(sizeof(int64_t) * (tg::kBins + 1) * (size_t)n_windows))) return rc;
    if ((rc = c->d_perm.reserve(sizeof(uint4) * (size_t)S))) return rc;
    int* d_nlong = (int*)c->d_qhead.p;
    int64_t* d_total = (int64_t*)((char*)c->d_qhead.p + 8);
    unsigned long long* d_ticket = (unsigned long long*)((char*)c->d_qhead.p + 16);
    int* d_long_grouped = (int*)((char*)c->d_qhead.p + 24);
    unsigned* d_hist = (unsigned*)c->d_hist.p;
    int64_t* d_start = (int64_t*)c->d_start.p;
    int64_t* d_wbase = d_start + tg::kBins * n_windows;          // per-window totals, then bases
    uint4* d_queue = (uint4*)c->d_perm.p;
    TG_CUDA(cudaMemsetAsync(c->d_qhead.p, 0, 32, st));          // n_long, total, ticket, long_grouped
    TG_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned) * tg::kBins * (size_t)n_windows, st));
    const unsigned seg_grid = (unsigned)((S + tg::kBinSeg - 1) / tg::kBinSeg);
    tg::k_bin_count<<<seg_grid, tg::kBinThreads, 0, st>>>(d_offsets, S, d_hist);
    tg::k_bin_scan<<<(unsigned)n_windows, 1024, 0, st>>>(d_hist, d_start, d_wbase);
    tg::k_window_scan<<<1, 1024, 0, st>>>(d_wbase, n_windows, d_total, d_long_grouped);
    tg::k_bin_scatter<<<seg_grid, tg::kBinThreads, 0, st>>>(d_offsets, S, d_hist, d_start, d_wbase, d_queue, d_out, ld, d_keep, d_nlong, d_long_grouped);
    const int64_t groups = (S + 31) / 32;
    const int64_t ctas = (groups + tg::kWarpsPerCta - 1) / tg::kWarpsPerCta;
    const unsigned grid = (unsigned)(ctas < c->sm_count ? ctas : c->sm_count);
    {
        tg::k_metrics_grouped<<<grid, tg::kGroupedThreads, tg::kGroupedSmem, st>>>(xyz, lo, hi, ld, d_queue, d_total, d_ticket, d_out, d_keep);
    }
    tg::k_metrics_long<<<(unsigned)c->sm_count * 4u, tg::kLongThreads, 0, st>>>(xyz, d_offsets, ld, (const unsigned*)(d_queue + S), d_nlong, d_out, d_keep);
    c->launches += 6;
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

static inline bool dtype_ok(int t) { return t == TG_F64 || t == TG_F32 || t == TG_F64_BE || t == TG_F32_BE; }
static inline size_t dtype_size(int t) { return (t == TG_F64 || t == TG_F64_BE) ? 8 : 4; }

// any storage form -> native float64 (count = number of coordinates)
static int decode_points(tg_context* c, const void* d_src, int xyz_dtype, int64_t count, double* d_dst, cudaStream_t st) {
    if (count <= 0) return TG_OK;
    const int64_t want = (count / 4 + 255) / 256;
    const int64_t cap = (int64_t)c->sm_count * 16;
    const unsigned g = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
    const int swap = (xyz_dtype == TG_F64_BE || xyz_dtype == TG_F32_BE) ? 1 : 0;
    if (dtype_size(xyz_dtype) == 4) tg::k_decode_points<true><<<g, 256, 0, st>>>(d_src, d_dst, count, swap);
    else tg::k_decode_points<false><<<g, 256, 0, st>>>(d_src, d_dst, count, swap);
    c->launches += 1;
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_metrics_csr_dev(tg_context* c, const void* d_xyz, int xyz_dtype, const int64_t* d_offsets, int64_t S, int64_t P,
                       double* d_out, uint8_t* d_keep, void* stream) {
    NvtxRange nvtx_("tg_metrics_csr_dev");
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (S < 0 || P < 0) return set_err(TG_E_INVALID, "negative size");
    if (!dtype_ok(xyz_dtype)) return set_err(TG_E_INVALID, "xyz_dtype must be TG_F64, TG_F32, TG_F64_BE or TG_F32_BE");
    if (S == 0) return TG_OK;
    if (!d_offsets || !d_out || !d_keep || (P > 0 && !d_xyz)) return set_err(TG_E_INVALID, "null device pointer");
    DeviceGuard g(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc;
    if (xyz_dtype != TG_F64) {
        if ((rc = c->d_xyz64.reserve(sizeof(double) * 3 * (size_t)P + 64))) return rc;
        if ((rc = decode_points(c, d_xyz, xyz_dtype, 3 * P, (double*)c->d_xyz64.p, st))) return rc;
        d_xyz = c->d_xyz64.p;
    }
    if (((uintptr_t)d_xyz & 7u) != 0) return set_err(TG_E_INVALID, "xyz must be 8-byte aligned");
    const uint64_t lo = (uint64_t)(uintptr_t)d_xyz;
    return launch_metrics_f64(c, (const double*)d_xyz, lo, lo + 24ull * (uint64_t)P, d_offsets, S, d_out, S, d_keep, st);
}

static int bundle_reduce_impl(tg_context* c, const double* d_out, const uint8_t* d_keep, const uint8_t* d_select, int64_t S,
                              const int64_t* h_bo, int64_t B, double* d_sums, int64_t* d_counts, double* d_packed, void* stream) {
    NvtxRange nvtx_("tg_bundle_reduce");
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (S < 0 || B < 0) return set_err(TG_E_INVALID, "negative size");
    if (B == 0) return TG_OK;
    if (!h_bo || (!d_packed && (!d_sums || !d_counts)) || (S > 0 && (!d_out || !d_keep))) return set_err(TG_E_INVALID, "null pointer");
    if (h_bo[0] < 0 || h_bo[B] > S) return set_err(TG_E_INVALID, "bundle_offsets out of range");
    for (int64_t b = 0; b < B; ++b)
        if (h_bo[b + 1] < h_bo[b]) return set_err(TG_E_INVALID, "bundle_offsets must be non-decreasing");
    DeviceGuard g(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = upload_bundle_src();
    if (rc) return rc;

    int64_t nt = 0;
    const tg::TileDesc* dt = nullptr;
    const int64_t* df = nullptr;
    if ((rc = plan_bundle_tiles(c, h_bo, B, st, &nt, &dt, &df))) return rc;
    if ((rc = c->d_tsum.reserve(sizeof(double) * tg::kNB * (size_t)(nt + 1)))) return rc;
    if ((rc = c->d_tcnt.reserve(sizeof(int64_t) * (tg::kNB + 1) * (size_t)(nt + 1)))) return rc;
    if (nt > 0) {
        tg::k_bundle_tiles<<<(unsigned)nt, tg::kBundleThreads, 0, st>>>(d_out, d_keep, d_select, S, dt, (double*)c->d_tsum.p, (int64_t*)c->d_tcnt.p);
        c->launches += 1;
    }
    tg::k_bundle_final<<<(unsigned)B, tg::kFinalThreads, 0, st>>>(df, (const double*)c->d_tsum.p, (const int64_t*)c->d_tcnt.p, d_sums, d_counts, d_packed);
    c->launches += 1;
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_bundle_reduce_dev(tg_context* c, const double* d_out, const uint8_t* d_keep, const uint8_t* d_select, int64_t S,
                         const int64_t* h_bo, int64_t B, double* d_sums, int64_t* d_counts, void* stream) {
    return bundle_reduce_impl(c, d_out, d_keep, d_select, S, h_bo, B, d_sums, d_counts, nullptr, stream);
}

int tg_bundle_partials_dev(tg_context* c, const double* d_out, const uint8_t* d_keep, const uint8_t* d_select, int64_t S,
                           const int64_t* h_bo, int64_t B, double* d_partials, void* stream) {
    if (!d_partials && B > 0) return set_err(TG_E_INVALID, "null pointer");
    return bundle_reduce_impl(c, d_out, d_keep, d_select, S, h_bo, B, nullptr, nullptr, d_partials, stream);
}

int tg_bundle_spread_dev(tg_context* c, const double* d_out, const uint8_t* d_keep, const uint8_t* d_select, int64_t S,
                         const int64_t* h_bo, int64_t B, const double* d_sums, const int64_t* d_counts, double* d_spread, void* stream) {
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (S < 0 || B < 0) return set_err(TG_E_INVALID, "negative size");
    if (B == 0) return TG_OK;
    if (!h_bo || !d_sums || !d_counts || !d_spread || (S > 0 && (!d_out || !d_keep))) return set_err(TG_E_INVALID, "null pointer");
    if (h_bo[0] < 0 || h_bo[B] > S) return set_err(TG_E_INVALID, "bundle_offsets out of range");
    for (int64_t b = 0; b < B; ++b)
        if (h_bo[b + 1] < h_bo[b]) return set_err(TG_E_INVALID, "bundle_offsets must be non-decreasing");
    DeviceGuard g(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = upload_bundle_src();
    if (rc) return rc;
    int64_t nt = 0;
    const tg::TileDesc* dt = nullptr;
    const int64_t* df = nullptr;
    if ((rc = plan_bundle_tiles(c, h_bo, B, st, &nt, &dt, &df))) return rc;
    if ((rc = c->d_tspread.reserve(sizeof(double) * 3 * tg::kNB * (size_t)(nt + 1)))) return rc;
    if (nt > 0) {
        tg::k_spread_tiles<<<(unsigned)nt, tg::kBundleThreads, 0, st>>>(d_out, d_keep, d_select, S, dt, d_sums, d_counts, (double*)c->d_tspread.p);
        c->launches += 1;
    }
    tg::k_spread_final<<<(unsigned)B, tg::kFinalThreads, 0, st>>>(df, (const double*)c->d_tspread.p, d_counts, d_spread);
    c->launches += 1;
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_resample_csr_dev(tg_context* c, const void* d_xyz, int xyz_dtype, const int64_t* d_offsets, int64_t S, int64_t P,
                        int n_nodes, double* d_nodes, void* stream) {
    NvtxRange nvtx_("tg_resample_csr_dev");
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (S < 0 || P < 0) return set_err(TG_E_INVALID, "negative size");
    if (n_nodes < 2) return set_err(TG_E_INVALID, "n_nodes must be at least 2");
    if (!dtype_ok(xyz_dtype)) return set_err(TG_E_INVALID, "xyz_dtype must be TG_F64, TG_F32, TG_F64_BE or TG_F32_BE");
    if (S == 0) return TG_OK;
    if (!d_offsets || !d_nodes || (P > 0 && !d_xyz)) return set_err(TG_E_INVALID, "null device pointer");
    DeviceGuard g(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc;
    if (xyz_dtype != TG_F64) {
        if ((rc = c->d_xyz64.reserve(sizeof(double) * 3 * (size_t)P + 64))) return rc;
        if ((rc = decode_points(c, d_xyz, xyz_dtype, 3 * P, (double*)c->d_xyz64.p, st))) return rc;
        d_xyz = c->d_xyz64.p;
    }
    if (((uintptr_t)d_xyz & 7u) != 0) return set_err(TG_E_INVALID, "xyz must be 8-byte aligned");
    if (!c->resample_ready) {
        TG_CUDA(cudaFuncSetAttribute(tg::k_resample, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::kResampleSmem));
        c->resample_ready = true;
    }
    const int64_t want = (S + tg::kRsWarps - 1) / tg::kRsWarps;
    const int64_t cap = (int64_t)c->sm_count * tg::kResampleCtasPerSm;
    const uint64_t lo = (uint64_t)(uintptr_t)d_xyz;
    tg::k_resample<<<(unsigned)(want < cap ? want : cap), tg::kResampleThreads, tg::kResampleSmem, st>>>(
        (const double*)d_xyz, lo, lo + 24ull * (uint64_t)P, d_offsets, S, n_nodes, d_nodes);
    c->launches += 1;
    TG_CUDA(cudaGetLastError());
    return TG_OK;
}

int tg_resample_csr_host(tg_context* c, const void* h_xyz, int xyz_dtype, const int64_t* h_off, int64_t S, int64_t P,
                         int n_nodes, double* h_nodes) {
    NvtxRange nvtx_("tg_resample_csr_host");
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (S < 0 || P < 0) return set_err(TG_E_INVALID, "negative size");
    if (n_nodes < 2) return set_err(TG_E_INVALID, "n_nodes must be at least 2");
    if (!dtype_ok(xyz_dtype)) return set_err(TG_E_INVALID, "xyz_dtype must be TG_F64, TG_F32, TG_F64_BE or TG_F32_BE");
    if (S == 0) return TG_OK;
    if (!h_off || !h_nodes || (P > 0 && !h_xyz)) return set_err(TG_E_INVALID, "null pointer");
    if (h_off[0] < 0 || h_off[S] > P) return set_err(TG_E_INVALID, "offsets out of range of the point array");
    for (int64_t s = 0; s < S; ++s)
        if (h_off[s + 1] < h_off[s]) return set_err(TG_E_INVALID, "offsets must be non-decreasing");
    DeviceGuard g(c->device);
    const size_t esz = dtype_size(xyz_dtype);
    const size_t node_bytes = sizeof(double) * 3 * (size_t)n_nodes * (size_t)S;
    int rc;
    if ((rc = c->d_xyz.reserve(esz * 3 * (size_t)P + 64))) return rc;
    if ((rc = c->d_off.reserve(sizeof(int64_t) * (size_t)(S + 1)))) return rc;
    if ((rc = c->d_nodes.reserve(node_bytes))) return rc;
    cudaStream_t st = c->stream;
    if (P > 0) TG_CUDA(cudaMemcpyAsync(c->d_xyz.p, h_xyz, esz * 3 * (size_t)P, cudaMemcpyHostToDevice, st));
    TG_CUDA(cudaMemcpyAsync(c->d_off.p, h_off, sizeof(int64_t) * (size_t)(S + 1), cudaMemcpyHostToDevice, st));
    if ((rc = tg_resample_csr_dev(c, c->d_xyz.p, xyz_dtype, (const int64_t*)c->d_off.p, S, P, n_nodes, (double*)c->d_nodes.p, st))) return rc;
    TG_CUDA(cudaMemcpyAsync(h_nodes, c->d_nodes.p, node_bytes, cudaMemcpyDeviceToHost, st));
    TG_CUDA(cudaStreamSynchronize(st));
    return TG_OK;
}

// ---- host-side ingest helpers: no device, no context --------------------------------------
int tg_vtk_lines_to_csr(const int64_t* lines, int64_t L, int64_t* offsets, int64_t* conn, int64_t* n_cells, int64_t* n_conn) {
    if (L < 0 || (L > 0 && !lines) || !offsets || !n_cells || !n_conn || (L > 0 && !conn)) return set_err(TG_E_INVALID, "null argument");
    int64_t i = 0, s = 0, c = 0;
    offsets[0] = 0;
    while (i < L) {
        const int64_t n = lines[i];
        if (n < 0 || n > L - i - 1) return set_err(TG_E_INVALID, "corrupt LINES array");
        for (int64_t k = 0; k < n; ++k) conn[c + k] = lines[i + 1 + k];
        c += n;
        offsets[++s] = c;
        i += 1 + n;
    }
    *n_cells = s;
    *n_conn = c;
    return TG_OK;
}

// The classic binary cell array as the file holds it (big-endian int32 `[n, i0..i(n-1), n, ...]`) -> CSR offsets in ONE
// pass, no widened copy; *identity = 1 when the connectivity is 0, 1, 2, ... (the usual tractography export: the points
// are used as they are and `conn` is not written), else conn (capacity L) receives the point indices.
int tg_vtk_cells_be32_to_csr(const void* cells_be, int64_t L, int64_t* offsets, int64_t* conn, int64_t* n_cells, int64_t* n_conn,
                             int* identity) {
    if (L < 0 || (L > 0 && !cells_be) || !offsets || !n_cells || !n_conn || !identity) return set_err(TG_E_INVALID, "null argument");
    const uint32_t* w = (const uint32_t*)cells_be;
    auto rd = [&](int64_t i) -> int64_t { return (int64_t)(int32_t)__builtin_bswap32(w[i]); };
    int64_t i = 0, s = 0, c = 0;
    bool ident = true;
    offsets[0] = 0;
    while (i < L) {
        const int64_t n = rd(i);
        if (n < 0 || n > L - i - 1) return set_err(TG_E_INVALID, "corrupt LINES array");
        if (ident) {
            for (int64_t k = 0; k < n; ++k)
                if (rd(i + 1 + k) != c + k) { ident = false; break; }
        }
        c += n;
        offsets[++s] = c;
        i += 1 + n;
    }
    if (!ident) {
        if (!conn) return set_err(TG_E_INVALID, "connectivity is not the identity and no buffer was given");
        i = 0; c = 0;
        while (i < L) {
            const int64_t n = rd(i);
            for (int64_t k = 0; k < n; ++k) conn[c + k] = rd(i + 1 + k);
            c += n;
            i += 1 + n;
        }
    }
    *n_cells = s;
    *n_conn = c;
    *identity = ident ? 1 : 0;
    return TG_OK;
}

static inline const char* skip_space(const char* p, const char* end) {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
    return p;
}

int tg_parse_ascii_f64(const char* text, int64_t len, int64_t want, double* out, int64_t* consumed) {
    if (len < 0 || want < 0 || (len > 0 && !text) || (want > 0 && !out) || !consumed) return set_err(TG_E_INVALID, "null argument");
    const char* p = text;
    const char* end = text + len;
    for (int64_t k = 0; k < want; ++k) {
        p = skip_space(p, end);
        if (p < end && *p == '+') ++p;                                  // from_chars takes no leading plus
        const std::from_chars_result r = std::from_chars(p, end, out[k]);
        if (r.ec == std::errc::result_out_of_range) {                   // strtod semantics: +-inf / 0 with the token consumed
            char tmp[64];
            const size_t m = std::min<size_t>(sizeof tmp - 1, (size_t)(r.ptr - p));
            memcpy(tmp, p, m); tmp[m] = 0;
            out[k] = strtod(tmp, nullptr);
        } else if (r.ec != std::errc()) {
            return set_err(TG_E_INVALID, p >= end ? "truncated ASCII block" : "malformed number in ASCII block");
        }
        p = r.ptr;
    }
    *consumed = (int64_t)(p - text);
    return TG_OK;
}

int tg_parse_ascii_i64(const char* text, int64_t len, int64_t want, int64_t* out, int64_t* consumed) {
    if (len < 0 || want < 0 || (len > 0 && !text) || (want > 0 && !out) || !consumed) return set_err(TG_E_INVALID, "null argument");
    const char* p = text;
    const char* end = text + len;
    for (int64_t k = 0; k < want; ++k) {
        p = skip_space(p, end);
        if (p < end && *p == '+') ++p;
        const std::from_chars_result r = std::from_chars(p, end, out[k]);
        if (r.ec != std::errc()) return set_err(TG_E_INVALID, p >= end ? "truncated ASCII block" : "malformed integer in ASCII block");
        p = r.ptr;
    }
    *consumed = (int64_t)(p - text);
    return TG_OK;
}

int tg_metrics_csr_host(tg_context* c, const void* h_xyz, int xyz_dtype, const int64_t* h_off, int64_t S, int64_t P,
                        const int64_t* h_bo, int64_t B, double* h_out, uint8_t* h_keep, double* h_sums, int64_t* h_counts) {
    return tg_metrics_csr_host_ex(c, h_xyz, xyz_dtype, h_off, S, P, h_bo, B, h_out, h_keep, h_sums, h_counts, nullptr);
}

static int host_pipeline(tg_context* c, const void* h_xyz, int xyz_dtype, const int64_t* h_off, int64_t S, int64_t P,
                         const int64_t* h_bo, int64_t B, double* h_out, uint8_t* h_keep, double* h_sums, int64_t* h_counts,
                         double* h_spread);
static int ensure_side_streams(tg_context* c);

// copies issued by a failed call may still be reading / writing the caller's buffers: drain before returning the error
static void drain_streams(tg_context* c) {
    if (c->s_copy) cudaStreamSynchronize(c->s_copy);
    cudaStreamSynchronize(c->stream);
    if (c->s_back) cudaStreamSynchronize(c->s_back);
    cudaGetLastError();
}

int tg_metrics_csr_host_ex(tg_context* c, const void* h_xyz, int xyz_dtype, const int64_t* h_off, int64_t S, int64_t P,
                           const int64_t* h_bo, int64_t B, double* h_out, uint8_t* h_keep, double* h_sums, int64_t* h_counts,
                           double* h_spread) {
    NvtxRange nvtx_("tg_metrics_csr_host_ex");
    if (!c) return set_err(TG_E_INVALID, "null context");
    const int rc = host_pipeline(c, h_xyz, xyz_dtype, h_off, S, P, h_bo, B, h_out, h_keep, h_sums, h_counts, h_spread);
    if (rc != TG_OK) drain_streams(c);
    return rc;
}

static int host_pipeline(tg_context* c, const void* h_xyz, int xyz_dtype, const int64_t* h_off, int64_t S, int64_t P,
                         const int64_t* h_bo, int64_t B, double* h_out, uint8_t* h_keep, double* h_sums, int64_t* h_counts,
                         double* h_spread) {
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (S < 0 || P < 0 || B < 0) return set_err(TG_E_INVALID, "negative size");
    if (!dtype_ok(xyz_dtype)) return set_err(TG_E_INVALID, "xyz_dtype must be TG_F64, TG_F32, TG_F64_BE or TG_F32_BE");
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
    const size_t esz = dtype_size(xyz_dtype);
    int rc;
    // ---- chunk plan: contiguous polyline ranges of at most ~chunk_points points each, so that the
    //      H2D copy of chunk i+1, the kernels of chunk i and the D2H of chunk i-1 overlap
    int64_t chunk_points = 8ll << 20;
    if (const char* e = getenv("TG_HOST_CHUNK_POINTS")) { long long v = atoll(e); if (v > 0) chunk_points = v; }
    std::vector<int64_t> cuts;
    cuts.push_back(0);
    int64_t max_pts = 0;
    while (cuts.back() < S) {
        const int64_t s0 = cuts.back();
        const int64_t* it = std::upper_bound(h_off + s0 + 1, h_off + S + 1, h_off[s0] + chunk_points);
        int64_t s1 = (int64_t)(it - h_off) - 1;
        if (s1 <= s0) s1 = s0 + 1;
        cuts.push_back(s1);
        max_pts = std::max(max_pts, h_off[s1] - h_off[s0]);
    }
    const int n_chunks = (int)cuts.size() - 1;
    // chunk buffers carry 256 bytes of slack on both sides: the staging of k_metrics_grouped may then read
    // whole sectors around the first / last polyline of a chunk, so no polyline is pushed to the exact path
    // just because of where the chunk boundary fell (results do not depend on the chunking)
    const size_t buf_stride = ((esz * 3 * (size_t)max_pts + 255) & ~(size_t)255) + 512;
    if ((rc = c->d_xyz.reserve(2 * buf_stride))) return rc;
    if (xyz_dtype != TG_F64 && (rc = c->d_xyz64.reserve(sizeof(double) * 3 * (size_t)max_pts + 512))) return rc;
    if ((rc = c->d_off.reserve(sizeof(int64_t) * (size_t)(S + 1)))) return rc;
    if ((rc = c->d_out.reserve(sizeof(double) * TG_N_METRICS * (size_t)S))) return rc;
    if ((rc = c->d_keep.reserve((size_t)S))) return rc;
    if ((rc = c->d_sums.reserve(sizeof(double) * tg::kNB * (size_t)B))) return rc;
    if ((rc = c->d_counts.reserve(sizeof(int64_t) * (tg::kNB + 1) * (size_t)B))) return rc;
    if (h_spread && (rc = c->d_spread.reserve(sizeof(double) * 3 * tg::kNB * (size_t)B))) return rc;
    if ((rc = ensure_side_streams(c))) return rc;
    cudaStream_t st = c->stream;
    double* d_out = (double*)c->d_out.p;
    uint8_t* d_keep = (uint8_t*)c->d_keep.p;
    const int64_t* d_off = (const int64_t*)c->d_off.p;
    TG_CUDA(cudaMemcpyAsync(c->d_off.p, h_off, sizeof(int64_t) * (size_t)(S + 1), cudaMemcpyHostToDevice, st));
    for (int i = 0; i < n_chunks; ++i) {
        const int b = i & 1;
        const int64_t s0 = cuts[i], s1 = cuts[i + 1], p0 = h_off[s0], np = h_off[s1] - p0;
        char* d_buf = (char*)c->d_xyz.p + (size_t)b * buf_stride + 256;
        if (i >= 2) TG_CUDA(cudaStreamWaitEvent(c->s_copy, c->ev_comp[b], 0));       // buffer b is free again
        if (np > 0) TG_CUDA(cudaMemcpyAsync(d_buf, (const char*)h_xyz + esz * 3 * (size_t)p0, esz * 3 * (size_t)np, cudaMemcpyHostToDevice, c->s_copy));
        TG_CUDA(cudaEventRecord(c->ev_h2d[b], c->s_copy));
        TG_CUDA(cudaStreamWaitEvent(st, c->ev_h2d[b], 0));
        const double* d_pts = (const double*)d_buf;
        if (xyz_dtype != TG_F64) {
            if ((rc = decode_points(c, d_buf, xyz_dtype, 3 * np, (double*)((char*)c->d_xyz64.p + 256), st))) return rc;
            d_pts = (const double*)((char*)c->d_xyz64.p + 256);
        }
        const uint64_t lo = (uint64_t)(uintptr_t)d_pts;
        const double* xyz_virtual = (const double*)(uintptr_t)(lo - 24ull * (uint64_t)p0);   // polyline s starts at xyz_virtual + 3*offsets[s]
        if ((rc = launch_metrics_f64(c, xyz_virtual, lo - 32, lo + 24ull * (uint64_t)np + 32, d_off + s0, s1 - s0, d_out + s0, S, d_keep + s0, st))) return rc;
        TG_CUDA(cudaEventRecord(c->ev_comp[b], st));
        if (h_out || h_keep) {
            TG_CUDA(cudaStreamWaitEvent(c->s_back, c->ev_comp[b], 0));
            if (h_out)      // the chunk's 17 column slices in ONE strided copy (pitch = a whole column)
                TG_CUDA(cudaMemcpy2DAsync(h_out + s0, sizeof(double) * (size_t)S, d_out + s0, sizeof(double) * (size_t)S,
                                          sizeof(double) * (size_t)(s1 - s0), TG_N_METRICS, cudaMemcpyDeviceToHost, c->s_back));
            if (h_keep) TG_CUDA(cudaMemcpyAsync(h_keep + s0, d_keep + s0, (size_t)(s1 - s0), cudaMemcpyDeviceToHost, c->s_back));
        }
    }
    if (B > 0) {
        if ((rc = tg_bundle_reduce_dev(c, d_out, d_keep, nullptr, S, h_bo, B, (double*)c->d_sums.p, (int64_t*)c->d_counts.p, st))) return rc;
        TG_CUDA(cudaMemcpyAsync(h_sums, c->d_sums.p, sizeof(double) * tg::kNB * (size_t)B, cudaMemcpyDeviceToHost, st));
        TG_CUDA(cudaMemcpyAsync(h_counts, c->d_counts.p, sizeof(int64_t) * (tg::kNB + 1) * (size_t)B, cudaMemcpyDeviceToHost, st));
        if (h_spread) {
            if ((rc = tg_bundle_spread_dev(c, d_out, d_keep, nullptr, S, h_bo, B, (const double*)c->d_sums.p, (const int64_t*)c->d_counts.p,
                                           (double*)c->d_spread.p, st))) return rc;
            TG_CUDA(cudaMemcpyAsync(h_spread, c->d_spread.p, sizeof(double) * 3 * tg::kNB * (size_t)B, cudaMemcpyDeviceToHost, st));
        }
    }
    TG_CUDA(cudaStreamSynchronize(c->s_copy));
    TG_CUDA(cudaStreamSynchronize(st));
    TG_CUDA(cudaStreamSynchronize(c->s_back));
    return TG_OK;
}

// ---- batch of files: points pushed file by file, one device call at the end ----------------
static int ensure_side_streams(tg_context* c) {
    if (!c->s_copy) {
        TG_CUDA(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
        TG_CUDA(cudaStreamCreateWithFlags(&c->s_back, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            TG_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
            TG_CUDA(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
        }
    }
    if (!c->ev_push) TG_CUDA(cudaEventCreateWithFlags(&c->ev_push, cudaEventDisableTiming));
    return TG_OK;
}

// more room for a running batch: new device buffers, the pushed part copied over behind the pending transfers
static int batch_grow(tg_context* c, int64_t need_P, int64_t need_S) {
    if (need_S > c->b_Scap) {
        const int64_t cap = std::max<int64_t>(need_S, 2 * c->b_Scap);
        PinBuf nb;
        int rc = nb.reserve(sizeof(int64_t) * (size_t)(cap + 1));
        if (rc) return rc;
        memcpy(nb.p, c->h_boff.p, sizeof(int64_t) * (size_t)(c->b_S + 1));
        c->h_boff.release();
        c->h_boff = nb;
        c->b_Scap = cap;
    }
    if (need_P > c->b_Pcap) {
        const int64_t cap = std::max<int64_t>(need_P, 2 * c->b_Pcap);
        DevBuf nxyz;
        int rc;
        if ((rc = nxyz.reserve(24 * (size_t)cap + 1024))) return rc;
        cudaError_t e = cudaStreamSynchronize(c->s_copy);            // pushed copies have landed; decodes are ordered on the main stream
        if (e == cudaSuccess && c->b_P > 0)
            e = cudaMemcpyAsync((char*)nxyz.p + 256, (char*)c->d_bxyz.p + 256, 24 * (size_t)c->b_P, cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            nxyz.release();
            return set_err(TG_E_CUDA, "growing the batch failed: %s", cudaGetErrorString(e));
        }
        c->d_bxyz.release();
        c->d_bxyz = nxyz;
        c->b_Pcap = cap;
    }
    return TG_OK;
}

int tg_batch_begin(tg_context* c, int64_t P_cap, int64_t S_cap) {
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (P_cap < 0 || S_cap < 0) return set_err(TG_E_INVALID, "negative size");
    DeviceGuard g(c->device);
    int rc;
    if ((rc = ensure_side_streams(c))) return rc;
    TG_CUDA(cudaStreamSynchronize(c->stream));                       // a previous batch may still read the buffers
    if ((rc = c->d_braw.reserve(std::min<size_t>(24 * (size_t)P_cap + 512, (size_t)64 << 20)))) return rc;
    if ((rc = c->d_bxyz.reserve(24 * (size_t)P_cap + 1024))) return rc;
    if ((rc = c->h_boff.reserve(sizeof(int64_t) * (size_t)(S_cap + 1)))) return rc;
    // the scratch is grow-only: a new batch starts with whatever capacity earlier batches left behind
    P_cap = std::max<int64_t>(P_cap, ((int64_t)c->d_bxyz.cap - 1024) / 24);
    S_cap = std::max<int64_t>(S_cap, (int64_t)(c->h_boff.cap / sizeof(int64_t)) - 1);
    c->b_P = 0; c->b_S = 0; c->b_raw = 0; c->b_Pcap = P_cap; c->b_Scap = S_cap;
    ((int64_t*)c->h_boff.p)[0] = 0;
    return TG_OK;
}

int tg_batch_push(tg_context* c, const void* h_xyz, int xyz_dtype, int64_t P_i, const int64_t* h_off, int64_t S_i) {
    NvtxRange nvtx_("tg_batch_push");
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (c->b_Pcap < 0) return set_err(TG_E_INVALID, "tg_batch_push without tg_batch_begin");
    if (!dtype_ok(xyz_dtype)) return set_err(TG_E_INVALID, "xyz_dtype must be TG_F64, TG_F32, TG_F64_BE or TG_F32_BE");
    if (P_i < 0 || S_i < 0 || (S_i > 0 && !h_off) || (P_i > 0 && !h_xyz)) return set_err(TG_E_INVALID, "bad argument");
    if (S_i > 0 && h_off[S_i] >= 0 && h_off[S_i] < P_i) P_i = h_off[S_i];   // points behind the last polyline are not part of the table
    if (S_i == 0) P_i = 0;
    if (c->b_P + P_i > c->b_Pcap || c->b_S + S_i > c->b_Scap) {        // capacities are hints: grow by doubling, keeping what was pushed
        DeviceGuard g2(c->device);
        int rc = batch_grow(c, c->b_P + P_i, c->b_S + S_i);
        if (rc) return rc;
    }
    if (S_i > 0 && (h_off[0] != 0 || h_off[S_i] > P_i)) return set_err(TG_E_INVALID, "offsets must start at 0 and stay inside the point array");
    int64_t* bo = (int64_t*)c->h_boff.p;
    for (int64_t s = 0; s < S_i; ++s) {
        const int64_t n = h_off[s + 1] - h_off[s];
        if (n < 0) return set_err(TG_E_INVALID, "offsets must be non-decreasing");
        if (n > 0x7ffffff0LL) return set_err(TG_E_INVALID, "a polyline has more than 2^31-16 points");
        bo[c->b_S + s + 1] = c->b_P + h_off[s + 1];
    }
    DeviceGuard g(c->device);
    if (P_i > 0) {
        const size_t esz = dtype_size(xyz_dtype), bytes = esz * 3 * (size_t)P_i;
        double* dst = (double*)((char*)c->d_bxyz.p + 256) + 3 * c->b_P;
        if (xyz_dtype == TG_F64) {                                   // native: straight into place
            TG_CUDA(cudaMemcpyAsync(dst, h_xyz, bytes, cudaMemcpyHostToDevice, c->s_copy));
        } else {                                                     // raw bytes now, decode behind the copy on the main stream
            // raw staging = a bump allocator over a modest device region: when it is full every earlier block has been
            // (or is about to be) decoded — wait for that and start over
            c->b_raw = (c->b_raw + 15) & ~(size_t)15;
            if (c->b_raw + bytes > c->d_braw.cap) {
                TG_CUDA(cudaStreamSynchronize(c->s_copy));
                TG_CUDA(cudaStreamSynchronize(c->stream));
                c->b_raw = 0;
                int rc = c->d_braw.reserve(bytes);
                if (rc) return rc;
            }
            char* raw = (char*)c->d_braw.p + c->b_raw;
            TG_CUDA(cudaMemcpyAsync(raw, h_xyz, bytes, cudaMemcpyHostToDevice, c->s_copy));
            TG_CUDA(cudaEventRecord(c->ev_push, c->s_copy));
            TG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_push, 0));
            int rc = decode_points(c, raw, xyz_dtype, 3 * P_i, dst, c->stream);
            if (rc) return rc;
            c->b_raw += bytes;
        }
    }
    c->b_P += P_i; c->b_S += S_i;
    return TG_OK;
}

int tg_batch_run(tg_context* c, const int64_t* h_bo, int64_t B, double* h_out, uint8_t* h_keep, double* h_sums, int64_t* h_counts,
                 double* h_spread) {
    NvtxRange nvtx_("tg_batch_run");
    if (!c) return set_err(TG_E_INVALID, "null context");
    if (c->b_Pcap < 0) return set_err(TG_E_INVALID, "tg_batch_run without tg_batch_begin");
    if (B < 0 || (B > 0 && (!h_bo || !h_sums || !h_counts))) return set_err(TG_E_INVALID, "bad bundle argument");
    const int64_t S = c->b_S, P = c->b_P;
    c->b_Pcap = -1;                                                  // the batch is consumed whatever happens
    if (S == 0) return TG_OK;
    DeviceGuard g(c->device);
    cudaStream_t st = c->stream;
    auto body = [&]() -> int {
        int rc;
        if ((rc = c->d_off.reserve(sizeof(int64_t) * (size_t)(S + 1)))) return rc;
        if ((rc = c->d_out.reserve(sizeof(double) * TG_N_METRICS * (size_t)S))) return rc;
        if ((rc = c->d_keep.reserve((size_t)S))) return rc;
        if ((rc = c->d_sums.reserve(sizeof(double) * tg::kNB * (size_t)B))) return rc;
        if ((rc = c->d_counts.reserve(sizeof(int64_t) * (tg::kNB + 1) * (size_t)B))) return rc;
        if (h_spread && (rc = c->d_spread.reserve(sizeof(double) * 3 * tg::kNB * (size_t)B))) return rc;
        TG_CUDA(cudaMemcpyAsync(c->d_off.p, c->h_boff.p, sizeof(int64_t) * (size_t)(S + 1), cudaMemcpyHostToDevice, c->s_copy));
        TG_CUDA(cudaEventRecord(c->ev_push, c->s_copy));             // every pushed copy precedes this event on s_copy
        TG_CUDA(cudaStreamWaitEvent(st, c->ev_push, 0));
        const double* d_pts = (const double*)((char*)c->d_bxyz.p + 256);
        const uint64_t lo = (uint64_t)(uintptr_t)d_pts;
        double* d_out = (double*)c->d_out.p;
        uint8_t* d_keep = (uint8_t*)c->d_keep.p;
        if ((rc = launch_metrics_f64(c, d_pts, lo - 32, lo + 24ull * (uint64_t)P + 32, (const int64_t*)c->d_off.p, S, d_out, S, d_keep, st))) return rc;
        if (h_out) TG_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(double) * TG_N_METRICS * (size_t)S, cudaMemcpyDeviceToHost, st));
        if (h_keep) TG_CUDA(cudaMemcpyAsync(h_keep, d_keep, (size_t)S, cudaMemcpyDeviceToHost, st));
        if (B > 0) {
            if ((rc = tg_bundle_reduce_dev(c, d_out, d_keep, nullptr, S, h_bo, B, (double*)c->d_sums.p, (int64_t*)c->d_counts.p, st))) return rc;
            TG_CUDA(cudaMemcpyAsync(h_sums, c->d_sums.p, sizeof(double) * tg::kNB * (size_t)B, cudaMemcpyDeviceToHost, st));
            TG_CUDA(cudaMemcpyAsync(h_counts, c->d_counts.p, sizeof(int64_t) * (tg::kNB + 1) * (size_t)B, cudaMemcpyDeviceToHost, st));
            if (h_spread) {
                if ((rc = tg_bundle_spread_dev(c, d_out, d_keep, nullptr, S, h_bo, B, (const double*)c->d_sums.p, (const int64_t*)c->d_counts.p,
                                               (double*)c->d_spread.p, st))) return rc;
                TG_CUDA(cudaMemcpyAsync(h_spread, c->d_spread.p, sizeof(double) * 3 * tg::kNB * (size_t)B, cudaMemcpyDeviceToHost, st));
            }
        }
        TG_CUDA(cudaStreamSynchronize(st));
        return TG_OK;
    };
    const int rc = body();
    if (rc != TG_OK) drain_streams(c);
    return rc;
}

int tg_batch_size(tg_context* c, int64_t* S, int64_t* P) {
    if (!c || !S || !P) return set_err(TG_E_INVALID, "null argument");
    *S = c->b_S; *P = c->b_P;
    return TG_OK;
}

}  // extern "C"
