// tg_grouped.cuh — kernel 1 of the streamline-metrics path (sm_100a): k_metrics_grouped.
//
// Work decomposition:
//   * a persistent grid of one 8-warp CTA per SM; every WARP is an independent worker that pulls
//     tiles of kTile = 256 consecutive polylines;
//   * the warp sorts its tile by point count (bitonic sort in shared memory) — the length-binned
//     work queue — and walks it in GROUPS of 32 polylines of (nearly) equal length, one polyline
//     per lane, so the lanes of a warp run the same number of iterations (no ragged divergence);
//   * the points of the 32 polylines of a group are staged through a per-lane double-buffered
//     ring in shared memory with cp.async (16-byte pieces, 8 lanes per polyline so that every
//     global request is a coalesced 128-byte line), kChunk = 12 points per slot, the next slot in
//     flight while the current one is consumed;
//   * each lane streams its polyline ONCE through a register pipeline
//         point -> segment/angle -> velocity -> binormal/curvature -> torsion
//     and keeps the 26 running sums in fp64 registers; nothing is shuffled or re-read.
//
// fp64-pipe budget: the B200 issues ~60 fp64 lane-operations / SM / clock (tools/microbench.cu),
// and at 25.45 algorithmic bytes per point that roof sits BELOW the HBM roof for this metric set,
// so the arithmetic below is organised to need ~100 fp64-pipe instructions per point (v1: ~145):
//   - every power-of-two scale factor of np.gradient is folded out: the pipeline carries
//     Vs = 2 v and B = 8 b (exact scalings), so the interior needs no multiplications by 1/2;
//   - kappa = |B| / (|Vs| + 2e-12)^3 comes from ONE rsqrt of |Vs|^6 |B|^2; the additive epsilon is
//     applied as a first-order correction inside the Newton step;
//   - the bending angle is evaluated as 2 asin(|t' - t| / 2) (no cancellation, no clipping), with
//     the reference's +1e-12 normalisation bias restored analytically, the asin tail in fp32;
//   - sum |t|^2 of the angular dispersion is 1 - 2e-12/|d| analytically;
//   - the finite test of the loader filter rides on the centroid sums (a non-finite coordinate
//     makes them non-finite); only then is the polyline re-scanned exactly.
// The streaming step is BRANCH-FREE and speculative: it assumes well-conditioned geometry
// (segments and velocities of ordinary magnitude, non-collinear triples, turning angle < 29 deg)
// and keeps a sticky per-lane `ok` flag; a polyline whose flag drops is recomputed by the exact
// general pipeline of tg_device.cuh (stream_chunk) by its lane after the group has finished.
// Two instantiations: STEADY (all 32 lanes in the interior of their polylines: no predicates at
// all) and GENERIC (ends of the polylines: every stage masked by selects).
//
// `ref:LINE` = /root/reference/src/geometry/tract_geom_proc.py:LINE.
#pragma once
#include "tg_device.cuh"

namespace tg {

constexpr int kWarpsPerCta = 8;
constexpr int kGroupedThreads = kWarpsPerCta * 32;
constexpr int kTile = 256;                    // polylines per warp tile (8 groups of 32)
constexpr int kChunk = 12;                    // points per ring slot
constexpr int kSub = 3;                       // steps per unrolled sub-block (= rotation period of the pipeline registers)
constexpr int kChunkBytes = kChunk * 24;      // 288
constexpr int kSlotBytes = kChunkBytes + 16;  // 16-byte lead-in: a point never straddles two slots
constexpr int kRingStride = 2 * kSlotBytes + 16;   // 624 B per lane: 16-B aligned, 2-way bank conflicts at worst
constexpr int kPieces = kSlotBytes / 16;      // 19 cp.async pieces per slot
constexpr int kWarpSmem = 32 * kRingStride + 32 * 16 + kTile * 4;   // ring + descriptors + sort keys
constexpr int kGroupedSmem = kWarpsPerCta * kWarpSmem;
constexpr int kMaxGroupedN = 1 << 22;         // longer polylines are left to k_metrics_whole (flagged)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// x / 2^k for a normal double, on the integer pipe
__device__ __forceinline__ double scale_pow2_down(double x, int k) {
    return __hiloint2double(__double2hiint(x) - (k << 20), __double2loint(x));
}
__device__ __forceinline__ double sel(bool p, double a, double b) { return p ? a : b; }

// tail of asin(s)/s - 1 - z/6, z = s^2 <= 1/16, in fp32 (|tail| <= 3e-4, so fp32 leaves < 4e-11)
__device__ __forceinline__ float asin_tail_f32(float z) {
    float r = 0.01396484f;
    r = fmaf(r, z, 0.017352764f);
    r = fmaf(r, z, 0.022372159f);
    r = fmaf(r, z, 0.030381944f);
    r = fmaf(r, z, 0.044642857f);
    r = fmaf(r, z, 0.075f);
    return r * (z * z);
}

// ------------------------------------------------------------------------------------------
// per-lane state
// ------------------------------------------------------------------------------------------
struct Sums {
    double L, th;                  // sum |d_j|, sum theta_i
    double t0, t1, t2, su;         // sum t_j, sum (1 - |t_j|^2)/2
    double q0, q1, q2, q00, q01, q02, q11, q12, q22;   // moments of p - m
    double mn0, mn1, mn2, mx0, mx1, mx2;
    double kK, k1, k2, en, ta;     // curvature shift / shifted moments, energy, torsion sum
};
struct Pipe {
    double p1x, p1y, p1z, p2x, p2y, p2z;      // P(k-1), P(k-2)
    double tx, ty, tz, u_prev, len_prev;      // unit segment k-2, eps/|d_{k-2}|, |d_{k-2}|
    double vax, vay, vaz, vbx, vby, vbz;      // Vs_{k-2}, Vs_{k-3}       (Vs = 2 v)
    double bax, bay, baz, bba, bbx, bby, bbz; // B_{k-3}, |B_{k-3}|^2, B_{k-4}   (B = 8 b)
};

__device__ __forceinline__ void sums_init(Sums& A) {
    A.L = A.th = A.t0 = A.t1 = A.t2 = A.su = 0.0;
    A.q0 = A.q1 = A.q2 = A.q00 = A.q01 = A.q02 = A.q11 = A.q12 = A.q22 = 0.0;
    A.mn0 = A.mn1 = A.mn2 = __longlong_as_double(0x7ff0000000000000LL);
    A.mx0 = A.mx1 = A.mx2 = __longlong_as_double(0xfff0000000000000LL);
    A.kK = A.k1 = A.k2 = A.en = A.ta = 0.0;
}
__device__ __forceinline__ void pipe_init(Pipe& S) {
    S.p1x = S.p1y = S.p1z = S.p2x = S.p2y = S.p2z = 0.0;
    S.tx = S.ty = S.tz = S.u_prev = S.len_prev = 0.0;
    S.vax = S.vay = S.vaz = S.vbx = S.vby = S.vbz = 0.0;
    S.bax = S.bay = S.baz = S.bba = S.bbx = S.bby = S.bbz = 0.0;
}

constexpr int kHi_4em9 = 0x3E312DFE;    // high word of ~4e-9: |Vs|^2 above it <=> 2 eps/|Vs| < 3.3e-8 (second-order term < 1e-14)
constexpr int kHi_1em8 = 0x3E45798F;    // high word of ~1e-8: |d|^2 above it <=> eps/|d| < 1e-8
constexpr int kHi_quarter = 0x3FD00000; // high word of 0.25

// One speculative pipeline step: consume P(k) = (cx,cy,cz).
//   stages: point k | segment k-1, angle k-2 | Vs k-1 | B, curvature k-2 | torsion k-3
// GENERIC = false: the caller guarantees 4 <= k <= n-1 (every stage live, no end effects).
// GENERIC = true : any k (also k < 0 or k > n+2: no effect); the caller passes P(min(k, n-1)).
// `ok` stays true only while every shortcut below is valid for this lane's data.
template <bool GENERIC>
__device__ __forceinline__ void lane_step(const int k, const int n, const double cx, const double cy, const double cz,
                                          const double m0, const double m1, const double m2, Pipe& S, Sums& A, bool& ok) {
    const int last = n - 1;
    const bool pP = !GENERIC || (k >= 0 && k <= last);          // point stage live
    const bool pS = !GENERIC || (k >= 1 && k <= last);          // segment j = k-1
    const bool pA = !GENERIC || (k >= 2 && k <= last);          // angle i = k-2
    const bool pV = !GENERIC || (k >= 0 && k <= n);             // Vs_{k-1} is new (else hold Vs_{n-1})
    const bool pB = !GENERIC || (k >= 2 && k <= n + 1);         // B_{k-2} is new (else hold B_{n-1})
    const bool pE = !GENERIC || (k >= 2 && k <= n);             // energy term j = k-2 < last
    const bool pT = !GENERIC || (n >= 4 && k >= 3 && k <= n + 2);   // torsion j = k-3
    if (GENERIC && k == 0) { S.p1x = S.p2x = cx; S.p1y = S.p2y = cy; S.p1z = S.p2z = cz; }    // left clamp P(-1) := P(0)

    // ---- point stage (ref:111-121 moments about m; ref:114-117 bounding box)
    {
        double qx = cx - m0, qy = cy - m1, qz = cz - m2;
        if (GENERIC) { qx = sel(pP, qx, 0.0); qy = sel(pP, qy, 0.0); qz = sel(pP, qz, 0.0); }
        A.q0 += qx; A.q1 += qy; A.q2 += qz;
        A.q00 = fma(qx, qx, A.q00); A.q01 = fma(qx, qy, A.q01); A.q02 = fma(qx, qz, A.q02);
        A.q11 = fma(qy, qy, A.q11); A.q12 = fma(qy, qz, A.q12); A.q22 = fma(qz, qz, A.q22);
        A.mn0 = sel(pP && cx < A.mn0, cx, A.mn0); A.mx0 = sel(pP && cx > A.mx0, cx, A.mx0);
        A.mn1 = sel(pP && cy < A.mn1, cy, A.mn1); A.mx1 = sel(pP && cy > A.mx1, cy, A.mx1);
        A.mn2 = sel(pP && cz < A.mn2, cz, A.mn2); A.mx2 = sel(pP && cz > A.mx2, cz, A.mx2);
    }

    // ---- segment stage j = k-1 (ref:32-33, 102, 145) and bending angle i = k-2 (ref:98-106)
    double len_cur;
    {
        double dx = cx - S.p1x, dy = cy - S.p1y, dz = cz - S.p1z;
        double x2 = fma(dz, dz, fma(dy, dy, dx * dx));
        ok = ok && (!pS || hi_in_range(x2, kHi_1em8, kHi_1e300));
        double y = rsqrt_fast(x2);                   // 1/|d| to 2^-58
        len_cur = x2 * y;
        double u_cur = kEps * y;                     // eps/|d| < 1e-8
        double inv = fma(-u_cur, y, y);              // 1/(|d| + eps) up to (eps/|d|)^2
        double ux = dx * inv, uy = dy * inv, uz = dz * inv;
        // |t' - t|^2 / 4 + ((1-|t|^2) + (1-|t'|^2)) / 4 = (1 - t.t')/2 for the eps-biased unit vectors
        double ex = ux - S.tx, ey = uy - S.ty, ez = uz - S.tz;
        double dd = fma(ez, ez, fma(ey, ey, ex * ex));
        double Z = fma(2.0, u_cur + S.u_prev, dd);   // 4 sin^2(theta/2)
        ok = ok && (!pA || __double2hiint(Z) < kHi_quarter);
        double yz = mufu_rsqrt(Z);                   // quadratic step: 1.3e-12 relative on theta
        double tz_ = Z * yz;
        double ez_ = fma(-tz_, yz, 1.0);
        double sZ = fma(scale_pow2_down(tz_, 1), ez_, tz_);      // sqrt(Z) = 2 sin(theta/2)
        float zf = 0.25f * __double2float_rn(Z);
        double w = fma(Z, 1.0 / 24.0, (double)asin_tail_f32(zf));
        double theta = fma(sZ, w, sZ);               // 2 asin(sqrt(Z)/2)
        if (GENERIC) {
            A.L += sel(pS, len_cur, 0.0);
            A.su += sel(pS, u_cur, 0.0);
            A.t0 += sel(pS, ux, 0.0); A.t1 += sel(pS, uy, 0.0); A.t2 += sel(pS, uz, 0.0);
            A.th += sel(pA, theta, 0.0);
            S.tx = sel(pS, ux, S.tx); S.ty = sel(pS, uy, S.ty); S.tz = sel(pS, uz, S.tz);
            S.u_prev = sel(pS, u_cur, S.u_prev);
            len_cur = sel(pS, len_cur, 0.0);
        } else {
            A.L += len_cur;
            A.su += u_cur;
            A.t0 += ux; A.t1 += uy; A.t2 += uz;
            A.th += theta;
            S.tx = ux; S.ty = uy; S.tz = uz; S.u_prev = u_cur;
        }
    }

    // ---- velocity stage j = k-1:  Vs_j = 2 v_j = (2 s_j)(P(j+1) - P(j-1)), clamped indices (ref:49)
    double vnx = cx - S.p2x, vny = cy - S.p2y, vnz = cz - S.p2z;
    if (GENERIC) {
        const double f = (k == 1 || k == n) ? 2.0 : 1.0;                   // one-sided ends: s = 1
        vnx = sel(pV, vnx * f, S.vax); vny = sel(pV, vny * f, S.vay); vnz = sel(pV, vnz * f, S.vaz);
    }

    // ---- binormal / curvature stage j = k-2:  B_j = 8 b_j (ref:50, 57-59)
    double bnx, bny, bnz, bbn;
    {
        double ex = vnx - S.vbx, ey = vny - S.vby, ez = vnz - S.vbz;       // 2 (v_{j+1} - v_{j-1})
        bnx = fma(S.vay, ez, -(S.vaz * ey));                               // numpy cross order, SURVEY.md N2
        bny = fma(S.vaz, ex, -(S.vax * ez));
        bnz = fma(S.vax, ey, -(S.vay * ex));
        if (GENERIC) {
            const double f = (k == 2 || k == n + 1) ? 2.0 : 1.0;           // s = 1 at the ends
            bnx *= f; bny *= f; bnz *= f;
        }
        bbn = fma(bnz, bnz, fma(bny, bny, bnx * bnx));
        double vv = fma(S.vaz, S.vaz, fma(S.vay, S.vay, S.vax * S.vax));  // |Vs_j|^2 = 4 |v_j|^2
        double w = (vv * vv) * (vv * bbn);                                 // |Vs|^6 |B|^2
        ok = ok && (!pB || (hi_in_range(w, kHi_1em280, kHi_1e300) && hi_in_range(vv, kHi_4em9, kHi_1e300)));
        // kappa = |b| / (|v| + eps)^3 = |B| / (|Vs| + 2 eps)^3 = |B|^2 rsqrt(w) (1 - 6 eps/|Vs| + ...)
        double r = mufu_rsqrt(w);
        double yv = mufu_rsqrt(vv);                                        // ~1/|Vs|: 20 bits is plenty for the 1e-11 term
        double t = w * r;
        double e = fma(-t, r, 1.0);
        e = fma(-12.0 * kEps, yv, e);                                      // fold (1 - 6 eps yv) into the Newton update (x 1/2)
        r = fma(scale_pow2_down(r, 1), e, r);
        double kappa = bbn * r;                                            // ref:59; finite on the speculative path
        const bool first = GENERIC && (k == 2);
        if (first) A.kK = kappa;                                           // shift for the moments: kappa_0
        double dk = kappa - A.kK;
        double ek = (kappa * kappa) * (S.len_prev + kEps);                 // ref:77,82-83
        if (GENERIC) {
            dk = sel(pB, dk, 0.0);
            ek = sel(pE, ek, 0.0);
            bnx = sel(pB, bnx, S.bax); bny = sel(pB, bny, S.bay); bnz = sel(pB, bnz, S.baz); bbn = sel(pB, bbn, S.bba);
        }
        A.k1 += dk;
        A.k2 = fma(dk, dk, A.k2);
        A.en += ek;
    }

    // ---- torsion stage j = k-3:  tau = b.db/(|b|^2 + eps) = s_j B_j.(B_{j+1} - B_{j-1}) / (|B_j|^2 + 64 eps)   (ref:91-95)
    {
        double ex = bnx - S.bbx, ey = bny - S.bby, ez = bnz - S.bbz;
        double num = fma(S.baz, ez, fma(S.bay, ey, S.bax * ex));
        double den = S.bba + 64.0 * kEps;
        double rc = rcp_fast(den);                                         // den in [6.4e-11, 1e300) whenever ok
        if (GENERIC) {
            const double f = (k == 3 || k == n + 2) ? 1.0 : 0.5;
            A.ta += sel(pT, (num * f) * rc, 0.0);
        } else {
            A.ta = fma(num, scale_pow2_down(rc, 1), A.ta);
        }
    }

    // ---- rotate (left clamp F(-1) := F(0) for the two gradient levels)
    if (GENERIC) {
        const bool c1 = (k == 1), c2 = (k == 2);
        S.vbx = sel(c1, vnx, S.vax); S.vby = sel(c1, vny, S.vay); S.vbz = sel(c1, vnz, S.vaz);
        S.bbx = sel(c2, bnx, S.bax); S.bby = sel(c2, bny, S.bay); S.bbz = sel(c2, bnz, S.baz);
    } else {
        S.vbx = S.vax; S.vby = S.vay; S.vbz = S.vaz;
        S.bbx = S.bax; S.bby = S.bay; S.bbz = S.baz;
    }
    S.vax = vnx; S.vay = vny; S.vaz = vnz;
    S.bax = bnx; S.bay = bny; S.baz = bnz; S.bba = bbn;
    S.len_prev = len_cur;
    S.p2x = S.p1x; S.p2y = S.p1y; S.p2z = S.p1z;
    S.p1x = cx; S.p1y = cy; S.p1z = cz;
}

// exact loader test ref:21 for a polyline whose centroid sums came out non-finite
__device__ __noinline__ bool all_finite_scan(const double* __restrict__ base, int n) {
    bool ok = true;
    for (int i = 0; i < 3 * n; ++i) ok = ok && finite_d(__ldg(base + i));
    return ok;
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

// 17 metrics from the sums of a complete, well-conditioned polyline (n >= 3); column-major out.
__device__ __forceinline__ unsigned finalize_grouped(const Sums& A, const int n,
                                                     const double f0, const double f1, const double f2,
                                                     const double e0, const double e1, const double e2,
                                                     const double m0, const double m1, const double m2,
                                                     double* __restrict__ out, const int64_t S, const int64_t s) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double dn = (double)n;
    const double L = A.L;                                                  // > 1e-4 (n-1) on this path: ref:160 passes
    double cx = e0 - f0, cy = e1 - f1, cz = e2 - f2;
    double chord = sqrt(cx * cx + cy * cy + cz * cz);                      // ref:36
    out[0 * S + s] = L;
    out[1 * S + s] = chord;
    out[2 * S + s] = L / fmax(chord, kMinLen);                             // ref:38-41
    out[3 * S + s] = chord / fmax(L, kMinLen);                             // ref:43-46
    {                                                                      // ref:61,71: all n curvatures are finite here
        double dm = A.k1 / dn;
        double m2c = A.k2 - A.k1 * dm;
        out[4 * S + s] = A.kK + dm;
        out[5 * S + s] = sqrt(fmax(m2c, 0.0) / dn);
    }
    out[6 * S + s] = A.en;
    out[7 * S + s] = (n >= 4) ? A.ta / dn : 0.0;                           // ref:86-87,96
    out[8 * S + s] = A.th / (double)(n - 2);                               // ref:106
    out[9 * S + s] = ((A.mx0 - A.mn0) * (A.mx1 - A.mn1)) * (A.mx2 - A.mn2);   // ref:117
    double g0 = A.q0 / dn, g1 = A.q1 / dn, g2 = A.q2 / dn;                 // centroid - m
    double f = 1.0 / (dn - 1.0);                                           // np.cov scales by the reciprocal
    double c00 = fma(-A.q0, g0, A.q00) * f, c01 = fma(-A.q0, g1, A.q01) * f, c02 = fma(-A.q0, g2, A.q02) * f;
    double c11 = fma(-A.q1, g1, A.q11) * f, c12 = fma(-A.q1, g2, A.q12) * f, c22 = fma(-A.q2, g2, A.q22) * f;
    double l1, l2, l3;
    sym3_eigenvalues(c00, c01, c02, c11, c12, c22, l1, l2, l3);
    out[10 * S + s] = (l2 <= kEps) ? inf : l1 / l2;                        // ref:126-130
    out[11 * S + s] = (l3 <= kEps) ? inf : l2 / l3;                        // ref:132-136
    out[12 * S + s] = l1 / (((l1 + l2) + l3) + kEps);                      // ref:138-141
    out[13 * S + s] = m0 + g0;                                             // ref:183-185
    out[14 * S + s] = m1 + g1;
    out[15 * S + s] = m2 + g2;
    // ref:143-148: mean |t - tbar|^2 = mean |t|^2 - |tbar|^2, mean |t|^2 = 1 - 2 su/(n-1)
    double dm1 = (double)(n - 1);
    double a0 = A.t0 / dm1, a1 = A.t1 / dm1, a2 = A.t2 / dm1;
    double disp = (1.0 - (a0 * a0 + a1 * a1 + a2 * a2)) - 2.0 * A.su / dm1;
    out[16 * S + s] = fmax(disp, 0.0);
    return 3u;
}

// ------------------------------------------------------------------------------------------
// warp-level bitonic sort of kTile 32-bit keys in shared memory (ascending)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_sort_tile(uint32_t* __restrict__ a, const int lane) {
#pragma unroll 1
    for (int k = 2; k <= kTile; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int t = lane; t < kTile / 2; t += 32) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int p = i | j;
                uint32_t x = a[i], y = a[p];
                bool up = (i & k) == 0;
                if ((x > y) == up) { a[i] = y; a[p] = x; }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Kernel 1
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGroupedThreads, 1)
k_metrics_grouped(const double* __restrict__ xyz, const int64_t* __restrict__ offsets, const int64_t S,
                  double* __restrict__ out, uint8_t* __restrict__ keep, int* __restrict__ long_flag) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wsm = smem + warp * kWarpSmem;
    unsigned char* ring = wsm;                                        // 32 x kRingStride
    uint4* desc = (uint4*)(wsm + 32 * kRingStride);                   // per polyline: {src lo, src hi, total bytes, -}
    uint32_t* keys = (uint32_t*)(wsm + 32 * kRingStride + 32 * 16);   // kTile sort keys
    const uint32_t ring_u32 = smem_u32(ring);
    const unsigned char* my_ring = ring + lane * kRingStride;

    const int64_t n_tiles = (S + kTile - 1) / kTile;
    const int64_t warps_total = (int64_t)gridDim.x * kWarpsPerCta;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);

    for (int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + warp; tile < n_tiles; tile += warps_total) {
        const int64_t t0 = tile * kTile;
        const int cnt = (int)min((int64_t)kTile, S - t0);
        // ---- length-binned queue: sort the tile by point count
#pragma unroll
        for (int i = lane; i < kTile; i += 32) {
            uint32_t key = 0xffffffffu;
            if (i < cnt) {
                int64_t nn = __ldg(offsets + t0 + i + 1) - __ldg(offsets + t0 + i);
                uint32_t nk = (uint32_t)min(max(nn, (int64_t)0), (int64_t)0xffffff);
                key = (nk << 8) | (uint32_t)i;
            }
            keys[i] = key;
        }
        __syncwarp();
        warp_sort_tile(keys, lane);

        for (int g = 0; g < kTile / 32; ++g) {
            const uint32_t key = keys[g * 32 + lane];
            if (__all_sync(0xffffffffu, key == 0xffffffffu)) break;
            const bool valid = key != 0xffffffffu;
            const int64_t s = t0 + (int)(key & 0xffu);
            int64_t o0 = 0, o1 = 0;
            if (valid) { o0 = __ldg(offsets + s); o1 = __ldg(offsets + s + 1); }
            const int64_t n64 = o1 - o0;
            if (valid && n64 > kMaxGroupedN) atomicOr(long_flag, 1);
            const bool act = valid && n64 >= 3 && n64 <= kMaxGroupedN;    // ref:21  sl.shape[0] > 2
            if (valid && n64 < 3) {
#pragma unroll
                for (int m = 0; m < 17; ++m) out[(int64_t)m * S + s] = nan;
                keep[s] = 0;
            }
            const int n = act ? (int)n64 : 0;
            const double* base = xyz + 3 * o0;
            const uint64_t baddr = (uint64_t)(uintptr_t)base;
            const int skew = act ? (int)(baddr & 15u) : 0;                 // 0 or 8
            const int sk = skew >> 3;
            const uint64_t a0 = baddr - (uint64_t)skew;
            const uint32_t total = act ? (uint32_t)(skew + 24 * n) : 0u;   // bytes from a0 to the true end
            desc[lane] = make_uint4((uint32_t)a0, (uint32_t)(a0 >> 32), total, 0u);
            // global step index gi = k + sk, k = 0 .. n+2
            int gmax_w = act ? n + 2 + sk : -1;
            int nmin_w = act ? n : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                gmax_w = max(gmax_w, __shfl_xor_sync(0xffffffffu, gmax_w, o));
                nmin_w = min(nmin_w, __shfl_xor_sync(0xffffffffu, nmin_w, o));
            }
            if (gmax_w < 0) continue;
            const int rounds = gmax_w / kChunk + 1;
            __syncwarp();

            // shift point for the moments: the middle point (any finite constant works)
            double m0 = 0.0, m1 = 0.0, m2 = 0.0;
            if (act) {
                const double* pm = base + 3 * (int64_t)(n >> 1);
                m0 = __ldg(pm); m1 = __ldg(pm + 1); m2 = __ldg(pm + 2);
                if (!(finite_d(m0) && finite_d(m1) && finite_d(m2))) { m0 = m1 = m2 = 0.0; }
            }

            // cooperative stage of chunk q of all 32 polylines into slot q&1: 8 lanes per polyline
            auto stage_chunk = [&](const int q) {
                const int part = lane & 7;
                const int pos0 = q * kChunkBytes - 16 + part * 16;         // byte position in the aligned stream
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = 4 * i + (lane >> 3);
                    const uint4 d = desc[j];
                    const unsigned char* src = (const unsigned char*)(uintptr_t)(((uint64_t)d.y << 32) | (uint64_t)d.x);
                    const uint32_t dst = ring_u32 + j * kRingStride + (q & 1) * kSlotBytes + part * 16;
#pragma unroll
                    for (int mm = 0; mm < 3; ++mm) {
                        const int piece = part + 8 * mm;
                        const int pos = pos0 + 128 * mm;
                        const int rem = (int)d.z - pos;
                        if (piece < kPieces && pos >= 0 && rem > 0)
                            cp_async16(dst + 128 * mm, src + pos, min(rem, 16));
                    }
                }
                cp_async_commit();
            };

            Sums A;
            Pipe P;
            sums_init(A);
            pipe_init(P);
            bool ok = true;
            double cx = 0.0, cy = 0.0, cz = 0.0;
            stage_chunk(0);
#pragma unroll 1
            for (int q = 0; q < rounds; ++q) {
                if (q + 1 < rounds) stage_chunk(q + 1); else cp_async_commit();
                cp_async_wait<1>();
                __syncwarp();
                const unsigned char* slot = my_ring + (q & 1) * kSlotBytes + (sk ? 0 : 16);
#pragma unroll 1
                for (int b = 0; b < kChunk / kSub; ++b) {
                    const int k0 = q * kChunk + b * kSub - sk;
                    const double* pp = (const double*)(slot + 24 * kSub * b);
                    // steady <=> every lane is strictly inside its polyline for all kSub steps
                    const bool steady = !act || (k0 >= 4 && k0 + kSub - 1 <= n - 1);
                    if (__all_sync(0xffffffffu, steady)) {
#pragma unroll
                        for (int i = 0; i < kSub; ++i) {
                            cx = pp[3 * i]; cy = pp[3 * i + 1]; cz = pp[3 * i + 2];
                            lane_step<false>(k0 + i, n, cx, cy, cz, m0, m1, m2, P, A, ok);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < kSub; ++i) {
                            const int k = k0 + i;
                            const bool ld = act && k >= 0 && k < n;
                            cx = sel(ld, pp[3 * i], cx); cy = sel(ld, pp[3 * i + 1], cy); cz = sel(ld, pp[3 * i + 2], cz);
                            lane_step<true>(k, n, cx, cy, cz, m0, m1, m2, P, A, ok);
                        }
                    }
                }
                __syncwarp();
            }
            cp_async_wait<0>();

            if (act) {
                const bool fin = finite_d(A.q0) && finite_d(A.q1) && finite_d(A.q2);
                if (ok && fin) {
                    double f0 = __ldg(base), f1 = __ldg(base + 1), f2 = __ldg(base + 2);
                    const double* pe = base + 3 * (int64_t)(n - 1);
                    double e0 = __ldg(pe), e1 = __ldg(pe + 1), e2 = __ldg(pe + 2);
                    keep[s] = (uint8_t)finalize_grouped(A, n, f0, f1, f2, e0, e1, e2, m0, m1, m2, out, S, s);
                } else {
                    keep[s] = (uint8_t)slow_polyline(base, n, out, S, s);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace tg
