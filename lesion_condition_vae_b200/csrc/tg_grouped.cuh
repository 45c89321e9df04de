// tg_grouped.cuh — kernel 1 of the streamline-metrics path (sm_100a): the length-binned work
// queue (k_bin_count / k_bin_scan / k_bin_scatter) and the streaming kernel k_metrics_grouped.
//
// Work decomposition:
//   * QUEUE.  A device-side counting sort orders the polylines of every window of kWindow
//     consecutive polylines by point count (one bin per length, 3..kShortMax; shorter ones get their
//     NaN row at once).  Cutting that order into consecutive GROUPS of 32 gives groups whose members
//     have the same length (a few per cent straddle a bin boundary).  Polylines of up to
//     kMaxGroupedN points form one extra row at the head of the queue, longest first (so the longest
//     groups start first and never form the tail of the launch); beyond kMaxGroupedN they are flagged
//     for the one-warp-per-polyline kernel.  The route depends on the polyline's own length only.
//   * STREAM.  A persistent grid of one 8-warp CTA per SM; every warp is an independent worker
//     that takes groups round-robin, one polyline per lane, so all 32 lanes run the same number
//     of iterations — no ragged-length divergence, no halo recomputation, no shuffles.
//   * STAGING.  The points of the 32 polylines of a group go through a per-lane double-buffered
//     ring in shared memory, filled with cp.async (16-byte pieces, 8 lanes per polyline so every
//     global request is a whole 128-byte line), kChunk = 12 points per slot, the next slot in
//     flight while the current one is consumed.  Polylines start on 8-byte boundaries; the staged
//     stream starts on the 32-byte sector below (so every request is whole sectors) and a slot
//     carries one sector of lead-out, so a point never straddles two slots.
//   * PIPELINE.  Each lane streams its polyline ONCE through a register pipeline
//         point -> segment/angle -> velocity -> binormal/curvature -> torsion
//     and keeps the 26 running sums in fp64 registers.
//
// fp64-pipe budget: the B200 issues ~60 fp64 lane-operations / SM / clock (tools/microbench.cu),
// and at 25.45 algorithmic bytes per point that roof sits BELOW the HBM roof for this metric set,
// so the arithmetic is organised to need 96 fp64-pipe instructions per interior point (v1: ~145):
//   - every power-of-two scale factor of np.gradient is folded out: the pipeline carries
//     Vs = 2 v and B = 8 b (exact scalings), so the interior needs no multiplications by 1/2;
//   - kappa = |B| / (|Vs| + 2e-12)^3 comes from ONE rsqrt of |Vs|^6 |B|^2; the additive epsilon is
//     applied as a first-order correction inside the Newton step;
//   - the bending angle is evaluated as 2 asin(|t' - t| / 2) (no cancellation, no clipping), with
//     the reference's +1e-12 normalisation bias restored analytically, the asin tail in fp32;
//   - sum |t|^2 of the angular dispersion is 1 - 2e-12/|d| analytically;
//   - the finite test of the loader filter rides on the centroid sums (a non-finite coordinate
//     makes them non-finite).
// The streaming step is SPECULATIVE: it assumes well-conditioned geometry (segments and
// velocities of ordinary magnitude, non-collinear triples, turning angle < 60 degrees) and keeps
// a sticky per-lane `ok` flag; a polyline whose flag drops is recomputed by the exact general
// pipeline of tg_device.cuh (stream_chunk) by its lane after the group has finished.
// Three instantiations of the step share the same arithmetic, operation for operation (the
// library is built with -fmad=false so the compiler cannot contract them differently), hence a
// polyline's result does not depend on which group it landed in:
//   STEADY  all lanes strictly inside their polylines: no predicates at all;
//   EDGE    first/last steps of a group whose 32 polylines have the same length: warp-uniform branches;
//   MASKED  groups of mixed lengths: every stage masked per lane by selects.
//
// `ref:LINE` = /root/reference/src/geometry/tract_geom_proc.py:LINE.
#pragma once
#include "tg_device.cuh"

namespace tg {

#ifndef TG_STAGE_BULK
#define TG_STAGE_BULK 0      // 1: experiment — stage with per-lane TMA bulk copies (cp.async.bulk + mbarrier) instead of cp.async pieces
#endif
#ifndef TG_OUT_EVICT_LAST
#define TG_OUT_EVICT_LAST 1
#endif
#ifndef TG_WARPS
#define TG_WARPS 8
#endif
constexpr int kWarpsPerCta = TG_WARPS;
constexpr int kGroupedThreads = kWarpsPerCta * 32;
#ifndef TG_CHUNK
#define TG_CHUNK 12
#endif
#ifndef TG_SUB
#define TG_SUB 3
#endif
constexpr int kChunk = TG_CHUNK;              // points per ring slot (even, multiple of kSub)
constexpr int kSub = TG_SUB;                  // steps per unrolled sub-block (3 = rotation period of the pipeline registers)
constexpr int kHead = ((6 + kSub - 1) / kSub) * kSub;   // first steps of a polyline, run as specialised EDGE steps (>= 4 needed)
static_assert(kChunk % kSub == 0 && kChunk % 2 == 0 && kHead <= kChunk, "chunk / sub-block geometry");
constexpr int kChunkBytes = kChunk * 24;      // 288
constexpr int kSlotBytes = kChunkBytes + 32;  // one 32-byte sector of lead-out: a point never straddles two slots
constexpr int kRingStride = 2 * kSlotBytes + 16;   // 656 B per lane: 16-B aligned, 2-way bank conflicts at worst
constexpr int kPieces = kSlotBytes / 16;      // 20 cp.async pieces per slot
static_assert(kPieces <= 24, "stage_chunk issues at most 3 pieces per lane and polyline");
constexpr int kWarpSmem = 32 * kRingStride + 32 * 16 + (TG_STAGE_BULK ? 32 : 0);   // ring + stream descriptors (+ 2 mbarriers in the bulk-copy experiment)
constexpr int kGroupedSmem = kWarpsPerCta * kWarpSmem;

constexpr int kBins = 2048;                   // bins per queue row
// Queue rows.  Rows 1 .. n_windows: one per window of kWindow consecutive polylines, polylines of 3 .. kShortMax
// points, one bin per length, ascending.  Row 0 (first in the queue, so its groups start first): the polylines
// of kShortMax+1 .. kMaxGroupedN points of the WHOLE table, 2 lengths per bin, longest first.  Beyond kMaxGroupedN a
// polyline goes to the one-warp-per-polyline kernel.  The route of a polyline depends on ITS length only — never
// on how many others there are — so its 17 numbers are bit-identical however the table is sharded across GPUs
// or chunked by the host path (round 1 sent row 0 to the warp-per-polyline kernel when it held < 6144 polylines:
// faster for a handful of long polylines, but the two kernels round differently).
constexpr int kShortMax = 1024;
constexpr int kLongShift = 1;
constexpr int kMaxGroupedN = kShortMax + (kBins << kLongShift);   // 5120
__device__ __forceinline__ int long_bin(const int64_t n) { return kBins - 1 - (int)((n - (kShortMax + 1)) >> kLongShift); }
#ifndef TG_WINDOW_LOG2
#define TG_WINDOW_LOG2 17
#endif
constexpr int kWindowLog2 = TG_WINDOW_LOG2;
constexpr int64_t kWindow = (int64_t)1 << kWindowLog2;   // polylines per sorting window
constexpr int kBinSeg = 8192;                 // polylines per CTA of the queue kernels (divides kWindow)
constexpr int kBinThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 16-byte cp.async issued iff rem > LIM (one ISETP + one predicated LDGSTS)
// (L2 cache hint: see TG_IN_POLICY)
template <int LIM, int OFF>
__device__ __forceinline__ void cp_async16_if(uint32_t dst, const void* src, int rem, uint64_t policy) {
    asm volatile("{ .reg .pred p; setp.gt.s32 p, %2, %3; @p cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %4; }"
                 ::"r"(dst + OFF), "l"((const unsigned char*)src + OFF), "r"(rem), "n"(LIM), "l"(policy) : "memory");
}
#ifndef TG_IN_POLICY
// L2 policy of the point reads: 0 evict_first, 1 evict_normal, 2 evict_last.  A 128-byte line that a 288-byte slot
// fetch cuts in the middle is touched again one round later: with evict_first it has often left L2 by then and
// is filled from DRAM a second time (reads 1.19x the payload; 1.09x with 1 or 2).  The result lines then lose
// their privilege (partial write-backs: +0.3 GB per 4M polylines), the sum is 5.5 % less DRAM traffic.
#define TG_IN_POLICY 2
#endif
__device__ __forceinline__ uint64_t policy_point_reads() {
    uint64_t p;
#if TG_IN_POLICY == 1
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
#elif TG_IN_POLICY == 2
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
#else
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
#endif
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// result store that asks L2 to keep the line: the 8-byte stores of a window's polylines land in random
// order, and a line evicted before its four sectors are complete costs a partial write plus a fill
__device__ __forceinline__ void st_keep(double* p, double v, uint64_t policy) {
#if TG_OUT_EVICT_LAST
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy) : "memory");
#else
    *p = v;
#endif
}
// ticket counter: a plain atom.add in PTX — with atomicAdd() ptxas emits its warp-aggregation sequence, whose
// leader broadcast (SHFL) waits for the atomic right where it was issued
__device__ __forceinline__ unsigned long long take_ticket(unsigned long long* p) {
    unsigned long long old;
    asm volatile("atom.global.add.u64 %0, [%1], 1;" : "=l"(old) : "l"(p) : "memory");
    return old;
}
// bulk-copy (TMA, non-tensor) staging primitives of the TG_STAGE_BULK experiment (profiles/experiments/README.md)
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra W; }" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// x / 2^k for a normal double, on the integer pipe
__device__ __forceinline__ double scale_pow2_down(double x, int k) {
    return __hiloint2double(__double2hiint(x) - (k << 20), __double2loint(x));
}
__device__ __forceinline__ double sel(bool p, double a, double b) { return p ? a : b; }

// tail of asin(s)/s - 1 - z/6 - 3 z^2/40, z = s^2 <= 1/4 (turning angle <= 60 degrees), in fp32: z^3 R(z) with a
// degree-5 minimax R; |tail| <= 8.4e-4 and the fp32 evaluation is within 1.3e-10 of it over the whole range
__device__ __forceinline__ float asin_tail_f32(float z) {
    float r = 0.023272162f;
    r = fmaf(r, z, 0.010175252f);
    r = fmaf(r, z, 0.017879551f);
    r = fmaf(r, z, 0.022339951f);
    r = fmaf(r, z, 0.030382656f);
    r = fmaf(r, z, 0.044642854f);
    return r * (z * z * z);
}

// ==========================================================================================
// Queue kernels
// ==========================================================================================
// hist[(1 + w) * kBins + n]: polylines of window w with n points (3 <= n <= kShortMax);
// hist[long_bin(n)]: polylines of the whole table with kShortMax < n <= kMaxGroupedN points
__global__ void __launch_bounds__(kBinThreads)
k_bin_count(const int64_t* __restrict__ offsets, const int64_t S, unsigned* __restrict__ hist) {
    __shared__ unsigned sh[2 * kBins];       // [0, kBins): this window's row, [kBins, 2 kBins): row 0
    for (int i = threadIdx.x; i < 2 * kBins; i += kBinThreads) sh[i] = 0u;
    __syncthreads();
    const int64_t s0 = (int64_t)blockIdx.x * kBinSeg;
    const int64_t s1 = min(s0 + kBinSeg, S);
    for (int64_t s = s0 + threadIdx.x; s < s1; s += kBinThreads) {
        const int64_t n = __ldg(offsets + s + 1) - __ldg(offsets + s);
        if (n >= 3 && n <= kShortMax) atomicAdd(&sh[(int)n], 1u);
        else if (n > kShortMax && n <= kMaxGroupedN) atomicAdd(&sh[kBins + long_bin(n)], 1u);
    }
    __syncthreads();
    unsigned* h = hist + (1 + (s0 >> kWindowLog2)) * kBins;
    for (int i = threadIdx.x; i < kBins; i += kBinThreads) {
        if (sh[i] != 0u) atomicAdd(&h[i], sh[i]);
        if (sh[kBins + i] != 0u) atomicAdd(&hist[i], sh[kBins + i]);
    }
}

// hist[r] -> exclusive start positions inside row r (one CTA per row); wtotal[r] = polylines queued in row r.
// hist becomes the scatter cursor (zeroed).
__global__ void __launch_bounds__(1024)
k_bin_scan(unsigned* __restrict__ hist, int64_t* __restrict__ start, int64_t* __restrict__ wtotal) {
    __shared__ unsigned wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w = blockIdx.x;
    unsigned* h = hist + w * kBins;
    const unsigned a = h[2 * threadIdx.x], b = h[2 * threadIdx.x + 1];
    unsigned v = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned t = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        wsum[lane] = t;
    }
    __syncthreads();
    const int64_t base = (int64_t)(warp ? wsum[warp - 1] : 0u) + (v - (a + b));
    start[w * kBins + 2 * threadIdx.x] = base;
    start[w * kBins + 2 * threadIdx.x + 1] = base + a;
    h[2 * threadIdx.x] = 0u;
    h[2 * threadIdx.x + 1] = 0u;
    if (threadIdx.x == 0) wtotal[w] = (int64_t)wsum[31];
}

// wtotal -> exclusive row bases (in place); total[0] = queue length M.  One CTA; rows are few.
// Decides whether row 0 is worth queueing (long_grouped[0] = 1) or its polylines go to k_metrics_long.
__global__ void __launch_bounds__(1024)
k_window_scan(int64_t* __restrict__ wtotal, const int64_t n_windows, int64_t* __restrict__ total, int* __restrict__ long_grouped) {
    __shared__ int64_t wsum[32];
    __shared__ int64_t run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        run = 0;
        const bool grouped = true;                       // by length only (see kMaxGroupedN)
        long_grouped[0] = grouped ? 1 : 0;
        if (!grouped) wtotal[0] = 0;
    }
    __syncthreads();
    for (int64_t w0 = 0; w0 < n_windows; w0 += 1024) {
        const int64_t w = w0 + threadIdx.x;
        const int64_t a = w < n_windows ? wtotal[w] : 0;
        int64_t v = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int64_t t = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int64_t u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            wsum[lane] = t;
        }
        __syncthreads();
        if (w < n_windows) wtotal[w] = run + (warp ? wsum[warp - 1] : 0) + (v - a);
        __syncthreads();
        if (threadIdx.x == 0) run += wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = run;
}

// queue[row base + bin start + rank] = {offset lo, offset hi, n, polyline id}; polylines with n < 3 get their NaN
// row (ref:21); ids of the polylines left to k_metrics_long are appended, downwards, at the END of the
// queue buffer (records + ids never exceed its S x 16 bytes) and counted in n_long.
__global__ void __launch_bounds__(kBinThreads)
k_bin_scatter(const int64_t* __restrict__ offsets, const int64_t S, unsigned* __restrict__ cursor,
              const int64_t* __restrict__ start, const int64_t* __restrict__ wbase, uint4* __restrict__ queue, double* __restrict__ out,
              const int64_t ld, uint8_t* __restrict__ keep, int* __restrict__ n_long, const int* __restrict__ long_grouped) {
    __shared__ unsigned sh[2 * kBins];   // pass 1: count; then: next free rank inside this CTA's reservation
    for (int i = threadIdx.x; i < 2 * kBins; i += kBinThreads) sh[i] = 0u;
    __syncthreads();
    const int64_t s0 = (int64_t)blockIdx.x * kBinSeg;
    const int64_t s1 = min(s0 + kBinSeg, S);
    const int64_t row = 1 + (s0 >> kWindowLog2);
    const int64_t max_row0 = long_grouped[0] ? kMaxGroupedN : kShortMax;   // longest polyline that is queued
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    unsigned* long_end = (unsigned*)(queue + S);
    for (int64_t s = s0 + threadIdx.x; s < s1; s += kBinThreads) {
        const int64_t n = __ldg(offsets + s + 1) - __ldg(offsets + s);
        if (n >= 3 && n <= kShortMax) atomicAdd(&sh[(int)n], 1u);
        else if (n < 3) {
#pragma unroll
            for (int m = 0; m < 17; ++m) out[(int64_t)m * ld + s] = nan;
            keep[s] = 0;
        } else if (n <= max_row0) atomicAdd(&sh[kBins + long_bin(n)], 1u);
        else long_end[-1 - (int64_t)atomicAdd(n_long, 1)] = (unsigned)s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += kBinThreads) {
        const unsigned c = sh[i];
        if (c != 0u) sh[i] = atomicAdd(&cursor[row * kBins + i], c);     // reserve [r, r+c) in the bin
        const unsigned c0 = sh[kBins + i];
        if (c0 != 0u) sh[kBins + i] = atomicAdd(&cursor[i], c0);
    }
    __syncthreads();
    for (int64_t s = s0 + threadIdx.x; s < s1; s += kBinThreads) {
        const int64_t o0 = __ldg(offsets + s);
        const int64_t n = __ldg(offsets + s + 1) - o0;
        const uint4 record = make_uint4((unsigned)o0, (unsigned)((uint64_t)o0 >> 32), (unsigned)n, (unsigned)s);
        if (n >= 3 && n <= kShortMax) {
            const unsigned r = atomicAdd(&sh[(int)n], 1u);
            queue[wbase[row] + start[row * kBins + n] + r] = record;
        } else if (n > kShortMax && n <= max_row0) {
            const int b = long_bin(n);
            const unsigned r = atomicAdd(&sh[kBins + b], 1u);
            queue[wbase[0] + start[b] + r] = record;
        }
    }
}

// ==========================================================================================
// Per-lane state
// ==========================================================================================
struct Sums {
    double L, th;                  // sum |d_j|, sum theta_i
    double t0, t1, t2, su;         // sum t_j, sum (1 - |t_j|^2)/2
    double q0, q1, q2, q00, q01, q02, q11, q12, q22;   // moments of p - m
#ifndef TG_BBOX_INT
#define TG_BBOX_INT 0
#endif
#if TG_BBOX_INT
    long long mn0, mn1, mn2, mx0, mx1, mx2;   // bounding box as order-preserving integer keys (integer pipe, not fp64)
#else
    double mn0, mn1, mn2, mx0, mx1, mx2;
#endif
    double kK, k1, k2, en, ta;     // curvature shift / shifted moments, sum kappa_j^2 |d_j|, torsion sum
    double kl;                     // kappa_{n-1}: the energy's eps * sum_{j<n-1} kappa_j^2 term comes from the moments
};
struct Pipe {
    double p1x, p1y, p1z, p2x, p2y, p2z;      // P(k-1), P(k-2)
    double tx, ty, tz, u_prev, len_prev;      // unit segment k-2, eps/|d_{k-2}|, |d_{k-2}|
    double vax, vay, vaz, vbx, vby, vbz;      // Vs_{k-2}, Vs_{k-3}       (Vs = 2 v)
    double bax, bay, baz, bba, dprev;         // B_{k-3}, |B_{k-3}|^2, B_{k-4}.B_{k-3}   (B = 8 b)
};

__device__ __forceinline__ void sums_init(Sums& A) {
    A.L = A.th = A.t0 = A.t1 = A.t2 = A.su = 0.0;
    A.q0 = A.q1 = A.q2 = A.q00 = A.q01 = A.q02 = A.q11 = A.q12 = A.q22 = 0.0;
#if TG_BBOX_INT
    A.mn0 = A.mn1 = A.mn2 = 0x7fffffffffffffffLL;
    A.mx0 = A.mx1 = A.mx2 = (long long)0x8000000000000000ULL;
#else
    A.mn0 = A.mn1 = A.mn2 = __longlong_as_double(0x7ff0000000000000LL);
    A.mx0 = A.mx1 = A.mx2 = __longlong_as_double(0xfff0000000000000LL);
#endif
    A.kK = A.k1 = A.k2 = A.en = A.ta = A.kl = 0.0;
}
// double <-> signed 64-bit key with the same ordering (an involution: flip the low 63 bits of negatives)
__device__ __forceinline__ long long order_key(double x) {
    const long long b = __double_as_longlong(x);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double key_value(long long k) { return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL)); }
__device__ __forceinline__ double key_value(double k) { return k; }
__device__ __forceinline__ void pipe_init(Pipe& S) {
    S.p1x = S.p1y = S.p1z = S.p2x = S.p2y = S.p2z = 0.0;
    S.tx = S.ty = S.tz = S.u_prev = S.len_prev = 0.0;
    S.vax = S.vay = S.vaz = S.vbx = S.vby = S.vbz = 0.0;
    S.bax = S.bay = S.baz = S.bba = S.dprev = 0.0;
}

constexpr int kHi_4em9 = 0x3E312DFE;    // high word of ~4e-9: |Vs|^2 above it <=> 2 eps/|Vs| < 3.3e-8 (second-order term < 1e-14)
constexpr int kHi_1em8 = 0x3E45798F;    // high word of ~1e-8: |d|^2 above it <=> eps/|d| < 1e-8
constexpr int kHi_one = 0x3FF00000;     // high word of 1.0
// Validity of the speculative path is tracked as ONE unsigned maximum per lane (a VIADDMNMX per test):
// test i contributes hi(x) - lo_i; the lane is fine iff the maximum stays below kChkSpan.  A value below
// its lower limit wraps to ~2^32; the common span makes the upper limits lo_i * 2^1023 (>= 9e27).
constexpr unsigned kChkSpan = (unsigned)(kHi_1e300 - kHi_1em8);
__device__ __forceinline__ unsigned chk_range(double x, int lo_hi) { return (unsigned)(__double2hiint(x) - lo_hi); }
__device__ __forceinline__ unsigned chk_below(double x, int lim_hi) { return (unsigned)(__double2hiint(x) - lim_hi) + kChkSpan; }   // x > 0 and hi(x) < lim

enum StepMode { STEADY = 0, MASKED = 1, EDGE = 2 };

// ---- stage arithmetic, shared by the three modes (operation for operation) -------------------
struct SegOut { double len, u, ux, uy, uz; unsigned chk; };
__device__ __forceinline__ SegOut seg_math(const double dx, const double dy, const double dz) {
    SegOut o;
    const double x2 = fma(dz, dz, fma(dy, dy, dx * dx));
    o.chk = chk_range(x2, kHi_1em8);
    const double y = rsqrt_fast(x2);                 // 1/|d| to 2^-58
    o.len = x2 * y;
    o.u = kEps * y;                                  // eps/|d| < 1e-8
    const double inv = fma(-o.u, y, y);              // 1/(|d| + eps) up to (eps/|d|)^2     ref:102,145
    o.ux = dx * inv; o.uy = dy * inv; o.uz = dz * inv;
    return o;
}
// theta = arccos(clip(t.t')) of the eps-biased unit vectors (ref:103-105), as 2 asin(sqrt(Z)/2) with
// Z/4 = |t'-t|^2/4 + ((1-|t|^2) + (1-|t'|^2))/4 = (1 - t.t')/2
__device__ __forceinline__ double angle_math(const SegOut& s, const double tx, const double ty, const double tz,
                                             const double u_prev, unsigned& chk) {
    const double ex = s.ux - tx, ey = s.uy - ty, ez = s.uz - tz;
    const double dd = fma(ez, ez, fma(ey, ey, ex * ex));
    const double Z = fma(2.0, s.u + u_prev, dd);     // 4 sin^2(theta/2)
    chk = chk_below(Z, kHi_one);                     // Z = 4 sin^2(theta/2) < 1: turning angle < 60 degrees
    const double yz = mufu_rsqrt(Z);                 // quadratic step: 1.3e-12 relative on theta
    const double tz_ = Z * yz;
    const double ez_ = fma(-tz_, yz, 1.0);
    const double sZ = fma(scale_pow2_down(tz_, 1), ez_, tz_);      // sqrt(Z) = 2 sin(theta/2)
    const float zf = 0.25f * __double2float_rn(Z);
    const double w = fma(Z, fma(Z, 3.0 / 640.0, 1.0 / 24.0), (double)asin_tail_f32(zf));   // z/6 + 3 z^2/40 + tail, z = Z/4
    return fma(sZ, w, sZ);
}
// kappa_j = |b_j| / (|v_j| + eps)^3 = |B_j| / (|Vs_j| + 2 eps)^3 = |B|^2 rsqrt(|Vs|^6 |B|^2) (1 - 6 eps/|Vs| + ...)   ref:57-59
__device__ __forceinline__ double kappa_math(const double vax, const double vay, const double vaz, const double bbn, unsigned& chk) {
    const double vv = fma(vaz, vaz, fma(vay, vay, vax * vax));     // |Vs_j|^2 = 4 |v_j|^2
    const double w = (vv * vv) * (vv * bbn);                       // |Vs|^6 |B|^2
    chk = max(chk_range(w, kHi_1em280), chk_range(vv, kHi_4em9));
    double r = mufu_rsqrt(w);
    const double yv = mufu_rsqrt(vv);                              // ~1/|Vs|: 20 bits is plenty for the 1e-11 term
    const double t = w * r;
    double e = fma(-t, r, 1.0);
    e = fma(-12.0 * kEps, yv, e);                                  // fold (1 - 6 eps yv) into the Newton update (x 1/2)
    r = fma(scale_pow2_down(r, 1), e, r);
    return bbn * r;
}

// One speculative pipeline step: consume P(k) = (cx,cy,cz).
//   stages: point k | segment k-1, angle k-2 | Vs k-1 | B, curvature k-2 | torsion k-3
// STEADY: the caller guarantees 4 <= k <= n-1 (every stage live, no end effects).
// EDGE  : 0 <= k <= n+2, k and n warp-uniform; the caller passes P(min(k, n-1)).
// MASKED: any k (k < 0 or k > n+2: no effect); the caller passes P(clamp(k, 0, n-1)) once k >= 0.
// `bad` (see kChkSpan) stays below kChkSpan only while every shortcut is valid for this lane's data.
template <int MODE>
__device__ __forceinline__ void lane_step(const int k, const int n, const double cx, const double cy, const double cz,
                                          const double m0, const double m1, const double m2, Pipe& S, Sums& A, unsigned& bad) {
    const int last = n - 1;
    const bool pP = MODE == STEADY || (k >= 0 && k <= last);          // point stage live
    const bool pS = MODE == STEADY || (k >= 1 && k <= last);          // segment j = k-1
    const bool pA = MODE == STEADY || (k >= 2 && k <= last);          // angle i = k-2
    const bool pV = MODE == STEADY || (k >= 0 && k <= n);             // Vs_{k-1} is new (else hold Vs_{n-1})
    const bool pB = MODE == STEADY || (k >= 2 && k <= n + 1);         // B_{k-2} is new (else hold B_{n-1})
    const bool pE = MODE == STEADY || (k >= 2 && k <= n);             // energy term j = k-2 < last
    const bool pT = MODE == STEADY || (n >= 4 && k >= 3 && k <= n + 2);   // torsion j = k-3
    if (MODE != STEADY && k == 0) { S.p1x = S.p2x = cx; S.p1y = S.p2y = cy; S.p1z = S.p2z = cz; }    // left clamp P(-1) := P(0)

    // ---- point stage (ref:111-121 moments about m; ref:114-117 bounding box)
    if (MODE != EDGE || pP) {
        double qx = cx - m0, qy = cy - m1, qz = cz - m2;
        if (MODE == MASKED) { qx = sel(pP, qx, 0.0); qy = sel(pP, qy, 0.0); qz = sel(pP, qz, 0.0); }
        A.q0 += qx; A.q1 += qy; A.q2 += qz;
        A.q00 = fma(qx, qx, A.q00); A.q01 = fma(qx, qy, A.q01); A.q02 = fma(qx, qz, A.q02);
        A.q11 = fma(qy, qy, A.q11); A.q12 = fma(qy, qz, A.q12); A.q22 = fma(qz, qz, A.q22);
        const bool mp = MODE != MASKED || pP;
#if TG_BBOX_INT
        const long long kx = order_key(cx), ky = order_key(cy), kz = order_key(cz);
#else
        const double kx = cx, ky = cy, kz = cz;
#endif
        A.mn0 = (mp && kx < A.mn0) ? kx : A.mn0; A.mx0 = (mp && kx > A.mx0) ? kx : A.mx0;
        A.mn1 = (mp && ky < A.mn1) ? ky : A.mn1; A.mx1 = (mp && ky > A.mx1) ? ky : A.mx1;
        A.mn2 = (mp && kz < A.mn2) ? kz : A.mn2; A.mx2 = (mp && kz > A.mx2) ? kz : A.mx2;
    }

    // ---- segment stage j = k-1 (ref:32-33, 102, 145) and bending angle i = k-2 (ref:98-106)
    double len_cur = 0.0;
    if (MODE != EDGE || pS) {
        const SegOut s = seg_math(cx - S.p1x, cy - S.p1y, cz - S.p1z);
        bad = max(bad, (MODE == MASKED && !pS) ? 0u : s.chk);
        len_cur = s.len;
        if (MODE == MASKED) {
            A.L += sel(pS, s.len, 0.0);
            A.su += sel(pS, s.u, 0.0);
            A.t0 += sel(pS, s.ux, 0.0); A.t1 += sel(pS, s.uy, 0.0); A.t2 += sel(pS, s.uz, 0.0);
        } else {
            A.L += s.len;
            A.su += s.u;
            A.t0 += s.ux; A.t1 += s.uy; A.t2 += s.uz;
        }
        if (MODE != EDGE || pA) {
            unsigned achk;
            const double theta = angle_math(s, S.tx, S.ty, S.tz, S.u_prev, achk);
            bad = max(bad, (MODE == MASKED && !pA) ? 0u : achk);
            A.th += (MODE == MASKED) ? sel(pA, theta, 0.0) : theta;
        }
        S.tx = s.ux; S.ty = s.uy; S.tz = s.uz; S.u_prev = s.u;   // stale values past the ends are never read
    }

    // ---- velocity stage j = k-1:  Vs_j = 2 v_j = (2 s_j)(P(j+1) - P(j-1)), clamped indices (ref:49)
    double vnx = S.vax, vny = S.vay, vnz = S.vaz;                          // hold Vs_{n-1} once j >= n
    if (MODE != EDGE || pV) {
        double x = cx - S.p2x, y = cy - S.p2y, z = cz - S.p2z;
        if (MODE != STEADY) {
            const double f = (k == 1 || k == n) ? 2.0 : 1.0;               // one-sided ends: s = 1
            x *= f; y *= f; z *= f;
        }
        if (MODE == MASKED) { vnx = sel(pV, x, vnx); vny = sel(pV, y, vny); vnz = sel(pV, z, vnz); }
        else { vnx = x; vny = y; vnz = z; }
    }

    // ---- binormal / curvature stage j = k-2:  B_j = 8 b_j (ref:50, 57-59)
    double bnx = S.bax, bny = S.bay, bnz = S.baz, bbn = S.bba;             // hold B_{n-1} once j >= n
    if (MODE != EDGE || pB) {
        const double ex = vnx - S.vbx, ey = vny - S.vby, ez = vnz - S.vbz; // 2 (v_{j+1} - v_{j-1})
        double x = fma(S.vay, ez, -(S.vaz * ey));                          // numpy cross order, SURVEY.md N2
        double y = fma(S.vaz, ex, -(S.vax * ez));
        double z = fma(S.vax, ey, -(S.vay * ex));
        if (MODE != STEADY) {
            const double f = (k == 2 || k == n + 1) ? 2.0 : 1.0;           // s = 1 at the ends
            x *= f; y *= f; z *= f;
        }
        const double bb = fma(z, z, fma(y, y, x * x));
        unsigned kchk;
        const double kappa = kappa_math(S.vax, S.vay, S.vaz, bb, kchk);    // finite on the speculative path
        bad = max(bad, (MODE == MASKED && !pB) ? 0u : kchk);
        if (MODE != STEADY && k == 2) A.kK = kappa;                        // shift for the moments: kappa_0
        double dk = kappa - A.kK;
        double kk = kappa * kappa;                                         // ref:77,82-83: kappa^2 (|d_j| + eps); the eps part is added in finalize
        double ds = S.len_prev;
        if (MODE != STEADY) A.kl = (k == n + 1) ? kappa : A.kl;
        if (MODE == MASKED) {
            dk = sel(pB, dk, 0.0);
            kk = sel(pE, kk, 0.0); ds = sel(pE, ds, 0.0);                  // 0*0: a masked-off lane may hold NaN in either
            bnx = sel(pB, x, bnx); bny = sel(pB, y, bny); bnz = sel(pB, z, bnz); bbn = sel(pB, bb, bbn);
        } else {
            if (MODE == EDGE && !pE) kk = 0.0;
            bnx = x; bny = y; bnz = z; bbn = bb;
        }
        A.k1 += dk;
        A.k2 = fma(dk, dk, A.k2);
        A.en = fma(kk, ds, A.en);
    }

    // ---- torsion stage j = k-3:  tau = b.db/(|b|^2 + eps) = s_j B_j.(B_{j+1} - B_{j-1}) / (|B_j|^2 + 64 eps)   (ref:91-95)
    //      B_j.(B_{j+1} - B_{j-1}) = D_j - D_{j-1} with D_j = B_j.B_{j+1} (clamped ends: D_{-1} = |B_0|^2, D_{n-1} = |B_{n-1}|^2):
    //      one new dot product per step; the rounding error it adds to tau is ~1e-16 ABSOLUTE (tau = num/|B|^2)
    double dn = S.dprev;
    if (MODE != EDGE || pT) {
        dn = fma(S.baz, bnz, fma(S.bay, bny, S.bax * bnx));
        const double num = dn - S.dprev;
        const double den = S.bba + 64.0 * kEps;
        double rc = rcp_fast(den);                                         // den in [6.4e-11, 1e300) whenever the lane is valid
        const bool end = MODE != STEADY && (k == 3 || k == n + 2);         // s = 1 at the ends, 1/2 inside
        rc = end ? rc : scale_pow2_down(rc, 1);
        const double tau = num * rc;
        A.ta += (MODE == MASKED) ? sel(pT, tau, 0.0) : tau;
    }

    // ---- rotate (left clamp F(-1) := F(0) for the two gradient levels)
    if (MODE != STEADY) {
        const bool c1 = (k == 1);
        S.vbx = sel(c1, vnx, S.vax); S.vby = sel(c1, vny, S.vay); S.vbz = sel(c1, vnz, S.vaz);
        S.dprev = (k == 2) ? bbn : dn;                                     // D_{-1} = B_0.B_0
    } else {
        S.vbx = S.vax; S.vby = S.vay; S.vbz = S.vaz;
        S.dprev = dn;
    }
    S.vax = vnx; S.vay = vny; S.vaz = vnz;
    S.bax = bnx; S.bay = bny; S.baz = bnz; S.bba = bbn;
    S.len_prev = len_cur;
    S.p2x = S.p1x; S.p2y = S.p1y; S.p2z = S.p1z;
    S.p1x = cx; S.p1y = cy; S.p1z = cz;
}

// exact general pipeline (tg_device.cuh) for a polyline the speculative path rejected
__device__ __noinline__ unsigned slow_polyline(const double* __restrict__ base, const int n, double* __restrict__ out,
                                               const int64_t S, const int64_t s) {
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
    Acc A;
    acc_init(A);
    stream_chunk<double, true>(base, n, 0, n, rx, ry, rz, m0, m1, m2, A);
    return finalize_metrics(A, n, f0, f1, f2, e0, e1, e2, m0, m1, m2, out, S, s);
}

// the iterative eigen-solve as a call: rare on this path (sym3_eigenvalues_fast rejects only nearly isotropic or
// nearly flat covariances), and ~1000 instructions that need not sit in the streaming kernel's hot code
__device__ __noinline__ void sym3_eigenvalues_sweeps(double a00, double a01, double a02, double a11, double a12, double a22,
                                                     double& l1, double& l2, double& l3) {
    sym3_eigenvalues(a00, a01, a02, a11, a12, a22, l1, l2, l3);
}

// 17 metrics from the sums of a complete, well-conditioned polyline (n >= 3); column-major out.
__device__ __forceinline__ unsigned finalize_grouped(const Sums& A, const int n,
                                                     const double f0, const double f1, const double f2,
                                                     const double e0, const double e1, const double e2,
                                                     const double m0, const double m1, const double m2,
                                                     double* __restrict__ out, const int64_t S, const int64_t s, const uint64_t pol) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double dn = (double)n;
    const double rn = rcp_fast(dn), rn1 = rcp_fast(dn - 1.0);             // 1/n, 1/(n-1) to 2^-58
    const double L = A.L;                                                  // > 1e-4 (n-1) on this path: ref:160 passes
    double cx = e0 - f0, cy = e1 - f1, cz = e2 - f2;
    double chord = sqrt_fast(cx * cx + cy * cy + cz * cz);                 // ref:36 (divisions below: x * rcp_fast(y), 2^-58 relative)
    st_keep(out + 0 * S + s, L, pol);
    st_keep(out + 1 * S + s, chord, pol);
    st_keep(out + 2 * S + s, L * rcp_fast(fmax(chord, kMinLen)), pol);                             // ref:38-41
    st_keep(out + 3 * S + s, chord * rcp_fast(fmax(L, kMinLen)), pol);                             // ref:43-46
    {                                                                      // ref:61,71: all n curvatures are finite here
        double dm = A.k1 * rn;
        double m2c = A.k2 - A.k1 * dm;
        st_keep(out + 4 * S + s, A.kK + dm, pol);
        st_keep(out + 5 * S + s, sqrt_fast(fmax(m2c, 0.0) * rn), pol);
        // ref:82-83: sum_{j<n-1} kappa_j^2 (|d_j| + eps) = A.en + eps (sum_all kappa^2 - kappa_{n-1}^2), with
        // sum_all kappa^2 = k2 + 2 K k1 + n K^2 from the shifted moments (a 1e-12-relative term: any rounding is fine)
        const double sk2 = fma(A.kK, fma(dn, A.kK, 2.0 * A.k1), A.k2) - A.kl * A.kl;
        st_keep(out + 6 * S + s, fma(kEps, sk2, A.en), pol);
    }
        st_keep(out + 7 * S + s, (n >= 4) ? A.ta * rn : 0.0, pol);                           // ref:86-87,96
    st_keep(out + 8 * S + s, A.th * rcp_fast((double)(n - 2)), pol);                               // ref:106
    st_keep(out + 9 * S + s, ((key_value(A.mx0) - key_value(A.mn0)) * (key_value(A.mx1) - key_value(A.mn1))) * (key_value(A.mx2) - key_value(A.mn2)), pol);   // ref:117
    double g0 = A.q0 * rn, g1 = A.q1 * rn, g2 = A.q2 * rn;                 // centroid - m
    double c00 = fma(-A.q0, g0, A.q00) * rn1, c01 = fma(-A.q0, g1, A.q01) * rn1, c02 = fma(-A.q0, g2, A.q02) * rn1;
    double c11 = fma(-A.q1, g1, A.q11) * rn1, c12 = fma(-A.q1, g2, A.q12) * rn1, c22 = fma(-A.q2, g2, A.q22) * rn1;
    double l1, l2, l3;
    if (!sym3_eigenvalues_fast(c00, c01, c02, c11, c12, c22, l1, l2, l3)) sym3_eigenvalues_sweeps(c00, c01, c02, c11, c12, c22, l1, l2, l3);
    st_keep(out + 10 * S + s, (l2 <= kEps) ? inf : l1 * rcp_fast(l2), pol);                        // ref:126-130
    st_keep(out + 11 * S + s, (l3 <= kEps) ? inf : l2 * rcp_fast(l3), pol);                        // ref:132-136
    st_keep(out + 12 * S + s, l1 * rcp_fast(((l1 + l2) + l3) + kEps), pol);                      // ref:138-141
    st_keep(out + 13 * S + s, m0 + g0, pol);                                             // ref:183-185
    st_keep(out + 14 * S + s, m1 + g1, pol);
    st_keep(out + 15 * S + s, m2 + g2, pol);
    // ref:143-148: mean |t - tbar|^2 = mean |t|^2 - |tbar|^2, mean |t|^2 = 1 - 2 su/(n-1)
    double a0 = A.t0 * rn1, a1 = A.t1 * rn1, a2 = A.t2 * rn1;
    double disp = (1.0 - (a0 * a0 + a1 * a1 + a2 * a2)) - 2.0 * A.su * rn1;
    st_keep(out + 16 * S + s, fmax(disp, 0.0), pol);
    return 3u;
}

// ==========================================================================================
// Kernel 1
// ==========================================================================================
__global__ void __launch_bounds__(kGroupedThreads, 1)
// xyz may be a VIRTUAL base (chunked host path): [xyz_lo, xyz_hi) is the byte range that may be read;
// ld = column stride of `out` (polylines of the whole table).
k_metrics_grouped(const double* __restrict__ xyz, const uint64_t xyz_lo, const uint64_t xyz_hi, const int64_t ld,
                  const uint4* __restrict__ queue, const int64_t* __restrict__ queue_len, unsigned long long* __restrict__ ticket,
                  double* __restrict__ out, uint8_t* __restrict__ keep) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wsm = smem + warp * kWarpSmem;
    unsigned char* ring = wsm;                                        // 32 x kRingStride
    uint4* desc = (uint4*)(wsm + 32 * kRingStride);                   // per polyline: {src lo, src hi, staged bytes, -}
    const uint32_t ring_u32 = smem_u32(ring);
    const unsigned char* my_ring = ring + lane * kRingStride;
    const uint64_t xyz_end = xyz_hi;
    const uint64_t l2_stream = policy_point_reads();
    const uint64_t l2_keep = policy_evict_last();

#if TG_STAGE_BULK
    const uint32_t bar_u32 = smem_u32(wsm + 32 * kRingStride + 32 * 16);      // two mbarriers, one per ring slot, 32 arrivals each
    if (lane == 0) { mbar_init(bar_u32, 32); mbar_init(bar_u32 + 8, 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t bar_phase = 0u;                                                  // bit s = parity the next wait on slot s expects
#endif
    const int64_t M = *queue_len;
    const int64_t n_groups = (M + 31) >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * kWarpsPerCta;
    // staging role of this lane: pieces (lane & 7) + 8 mm of polylines 4 i + (lane >> 3)
    const int part = lane & 7;
    const uint32_t stage_dst0 = ring_u32 + (lane >> 3) * kRingStride + part * 16;
    const uint4* stage_desc = desc + (lane >> 3);

    // dynamic group queue: the first group of a warp is static, later ones come from a global ticket.
    // Two-deep software pipeline: the ticket for the group after next is taken while the record of the
    // next group is being fetched, so neither latency is exposed.
    int64_t g = (int64_t)blockIdx.x * kWarpsPerCta + warp;
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if (g < n_groups && (g << 5) + lane < M) rec = __ldg(queue + (g << 5) + lane);
    int64_t gn = 0;
    if (lane == 0) gn = warps_total + (int64_t)take_ticket(ticket);
    gn = __shfl_sync(0xffffffffu, gn, 0);

    while (g < n_groups) {
        const bool act = (g << 5) + lane < M;
        const int64_t o0 = (int64_t)(((uint64_t)rec.y << 32) | (uint64_t)rec.x);
        const int n = act ? (int)rec.z : 0;
        const int64_t s = (int64_t)rec.w;
        if (gn < n_groups && (gn << 5) + lane < M) rec = __ldg(queue + (gn << 5) + lane);   // used by the next iteration
        unsigned long long tk = 0ull;
        if (lane == 0) tk = take_ticket(ticket);                                          // used at the end of this one
        const double* base = xyz + 3 * o0;
        const uint64_t baddr = (uint64_t)(uintptr_t)base;
        const int skew = act ? (int)(baddr & 31u) : 0;                 // 0, 8, 16 or 24
        const uint64_t a0 = baddr - (uint64_t)skew;
        // bytes staged from a0: the polyline rounded up to whole 16-byte pieces — except when that
        // would read past the end of the point array (only the last polyline can): that one is
        // left to the exact path, which reads with plain 8-byte loads
        uint32_t total = act ? (uint32_t)((skew + 24 * n + 15) & ~15) : 0u;
        unsigned bad = 0u;
        if (act && a0 + total > xyz_end) { bad = ~0u; total -= 16u; }
        if (act && a0 < xyz_lo) { bad = ~0u; total = 0u; }   // would read below the array (unaligned xyz): exact path
        desc[lane] = make_uint4((uint32_t)a0, (uint32_t)(a0 >> 32), total, 0u);
        const int n0 = __shfl_sync(0xffffffffu, n, 0);
        const bool exact = __all_sync(0xffffffffu, act && n == n0) && n0 >= kHead + 2;
        int nmax = n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
        const int steps = nmax + 3;                                    // k = 0 .. nmax+2
        const int rounds = (steps + kChunk - 1) / kChunk;
        __syncwarp();

        // cooperative stage of chunk q of all 32 polylines into slot q&1: 8 lanes per polyline
        // Slot layout: [0,32) carry | [32,320) fresh.  Chunk 0 is staged whole; for q >= 1 only the 288 fresh
        // bytes come from memory (whole sectors, every byte fetched once) and the first sector is the previous
        // slot's last one, copied by the lane itself inside shared memory (carry_sector).
        // the 8 stream descriptors this lane stages from: fetched once per group, kept in registers
        uint64_t sd_src[8];
        int sd_len[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 d = stage_desc[4 * i];
            sd_src[i] = ((uint64_t)d.y << 32) | (uint64_t)d.x;
            sd_len[i] = (int)d.z;
        }
#if TG_STAGE_BULK
        // EXPERIMENT: every lane fetches the chunk of ITS OWN polyline with one TMA bulk copy that completes on the slot's
        // mbarrier.  Source and size are per-lane values, the instruction takes uniform registers: ptxas serialises it
        // over the lanes (profiles/experiments/README.md).
        const uint32_t my_ring_u32 = ring_u32 + lane * kRingStride;
        const uint64_t my_a0 = a0;
        const int my_total = (int)total;
        auto stage_chunk = [&](const int q) {
            const int skip = q > 0 ? 32 : 0;
            const int pos0 = q * kChunkBytes + skip;
            const int want = q > 0 ? kChunkBytes : kSlotBytes;
            const int bytes = min(max(my_total - pos0, 0), want);
            const uint32_t bar = bar_u32 + 8 * (q & 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the slot was last touched through the generic proxy
            if (bytes > 0) {
                mbar_arrive_tx(bar, (uint32_t)bytes);
                bulk_g2s(my_ring_u32 + (q & 1) * kSlotBytes + skip, (const void*)(uintptr_t)(my_a0 + (uint64_t)pos0), (uint32_t)bytes, bar, l2_stream);
            } else {
                mbar_arrive(bar);
            }
        };
        auto wait_chunk = [&](const int q) {
            mbar_wait(bar_u32 + 8 * (q & 1), (bar_phase >> (q & 1)) & 1u);
            bar_phase ^= 1u << (q & 1);
        };
#else
        auto stage_chunk = [&](const int q) {
            const int skip = q > 0 ? 32 : 0;
            const int pos0 = q * kChunkBytes + skip + part * 16;       // byte position in the aligned stream
            const uint32_t dst0 = stage_dst0 + (q & 1) * kSlotBytes + skip;
            const int pieces = q > 0 ? kChunkBytes / 16 : kPieces;     // 16-byte pieces of this slot that come from memory
            const bool second = part + 8 < pieces, third = part + 16 < pieces;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned char* src = (const unsigned char*)(uintptr_t)sd_src[i] + pos0;
                const int rem = sd_len[i] - pos0;
                const uint32_t dst = dst0 + i * (4 * kRingStride);
                cp_async16_if<0, 0>(dst, src, rem, l2_stream);
                if (kPieces > 16 || second) cp_async16_if<128, 128>(dst, src, rem, l2_stream);
                if (kPieces > 16 && third) cp_async16_if<256, 256>(dst, src, rem, l2_stream);
            }
            cp_async_commit();
        };
        auto wait_chunk = [&](const int) { cp_async_wait<1>(); };
#endif
        auto carry_sector = [&](const int q) {                         // end of round q: lead-out of slot q -> head of slot q+1
            const uint4* src = (const uint4*)(my_ring + (q & 1) * kSlotBytes + kChunkBytes);
            uint4* dst = (uint4*)(const_cast<unsigned char*>(my_ring) + ((q + 1) & 1) * kSlotBytes);
            const uint4 v0 = src[0], v1 = src[1];
            dst[0] = v0; dst[1] = v1;
        };

        Sums A;
        Pipe Q;
        sums_init(A);
        pipe_init(Q);
        double cx = 0.0, cy = 0.0, cz = 0.0;
        double m0 = 0.0, m1 = 0.0, m2 = 0.0;                           // shift of the moments: P(0), set at k = 0
        stage_chunk(0);
        if (exact) {
            // ---- all 32 polylines have n0 >= 8 points: every step is warp-uniform.
            //      head k = 0..5 and tail k = n0..n0+2 are compile-time specialisations of the EDGE step
            //      (every predicate folds), the interior runs the predicate-free STEADY step.
            __builtin_assume(n0 >= kHead + 2);
            const int rounds_e = (n0 + kChunk - 1) / kChunk;           // rounds that bring in points
#pragma unroll 1
            for (int q = 0; q < rounds_e; ++q) {
                if (q + 1 < rounds_e) stage_chunk(q + 1); else if (!TG_STAGE_BULK) cp_async_commit();
                wait_chunk(q);
                __syncwarp();
                const unsigned char* slot = my_ring + (q & 1) * kSlotBytes + skew;
                int b = 0;
                if (q == 0) {
                    const double* pp = (const double*)slot;
#pragma unroll
                    for (int k = 0; k < kHead; ++k) {
                        cx = pp[3 * k]; cy = pp[3 * k + 1]; cz = pp[3 * k + 2];
                        if (k == 0) { m0 = cx; m1 = cy; m2 = cz; }
                        lane_step<EDGE>(k, n0, cx, cy, cz, m0, m1, m2, Q, A, bad);
                    }
                    b = kHead / kSub;
                }
#pragma unroll 1
                for (; b < kChunk / kSub; ++b) {
                    const int k0 = q * kChunk + b * kSub;
                    const double* pp = (const double*)(slot + 24 * kSub * b);
                    if (k0 + kSub <= n0) {
#pragma unroll
                        for (int i = 0; i < kSub; ++i) {
                            cx = pp[3 * i]; cy = pp[3 * i + 1]; cz = pp[3 * i + 2];
                            lane_step<STEADY>(k0 + i, n0, cx, cy, cz, m0, m1, m2, Q, A, bad);
                        }
                    } else {
                        // the last 0..2 interior points
#pragma unroll 1
                        for (int k = k0; k < n0; ++k) {
                            cx = pp[3 * (k - k0)]; cy = pp[3 * (k - k0) + 1]; cz = pp[3 * (k - k0) + 2];
                            lane_step<STEADY>(k, n0, cx, cy, cz, m0, m1, m2, Q, A, bad);
                        }
                        break;
                    }
                }
                if (q + 1 < rounds_e) carry_sector(q);
                __syncwarp();
            }
            // drain: (cx,cy,cz) = P(n0-1) held
            lane_step<EDGE>(n0, n0, cx, cy, cz, m0, m1, m2, Q, A, bad);
            lane_step<EDGE>(n0 + 1, n0, cx, cy, cz, m0, m1, m2, Q, A, bad);
            lane_step<EDGE>(n0 + 2, n0, cx, cy, cz, m0, m1, m2, Q, A, bad);
        } else {
#pragma unroll 1
            for (int q = 0; q < rounds; ++q) {
                if (q + 1 < rounds) stage_chunk(q + 1); else if (!TG_STAGE_BULK) cp_async_commit();
                wait_chunk(q);
                __syncwarp();
                const unsigned char* slot = my_ring + (q & 1) * kSlotBytes + skew;
#pragma unroll 1
                for (int b = 0; b < kChunk / kSub; ++b) {
                    const int k0 = q * kChunk + b * kSub;
                    if (k0 >= steps) break;
                    const double* pp = (const double*)(slot + 24 * kSub * b);
                    // steady <=> every lane is strictly inside its polyline for all kSub steps
                    const bool steady = !act || (k0 >= 4 && k0 + kSub <= n);
                    if (__all_sync(0xffffffffu, steady)) {
#pragma unroll
                        for (int i = 0; i < kSub; ++i) {
                            cx = pp[3 * i]; cy = pp[3 * i + 1]; cz = pp[3 * i + 2];
                            lane_step<STEADY>(k0 + i, n, cx, cy, cz, m0, m1, m2, Q, A, bad);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < kSub; ++i) {
                            const int k = k0 + i;
                            const bool ld = act && k < n;
                            cx = sel(ld, pp[3 * i], cx); cy = sel(ld, pp[3 * i + 1], cy); cz = sel(ld, pp[3 * i + 2], cz);
                            if (k == 0) { m0 = cx; m1 = cy; m2 = cz; }
                            lane_step<MASKED>(k, n, cx, cy, cz, m0, m1, m2, Q, A, bad);
                        }
                    }
                }
                if (q + 1 < rounds) carry_sector(q);
                __syncwarp();
            }
        }
        if (!TG_STAGE_BULK) cp_async_wait<0>();

        if (act) {
            // (m0,m1,m2) = P(0) and (cx,cy,cz) = P(n-1) are still in registers
            const bool fin = finite_d(A.q0) && finite_d(A.q1) && finite_d(A.q2);
            if (bad < kChkSpan && fin) keep[s] = (uint8_t)finalize_grouped(A, n, m0, m1, m2, cx, cy, cz, m0, m1, m2, out, ld, s, l2_keep);
            else keep[s] = (uint8_t)slow_polyline(base, n, out, ld, s);
        }
        __syncwarp();
        g = gn;
        {   // broadcast the ticket only now: an opaque zero keeps the compiler from hoisting the shuffle up to the
            // atomic (where its latency would be exposed); asm volatile stays behind the cp.async statements above
            unsigned long long zero;
            asm volatile("mov.u64 %0, 0;" : "=l"(zero));
            gn = warps_total + (int64_t)__shfl_sync(0xffffffffu, tk + zero, 0);
        }
    }
}

}  // namespace tg
