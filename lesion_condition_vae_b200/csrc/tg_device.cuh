// tg_device.cuh — device-side arithmetic of the streamline-metrics path (sm_100a).
//
// What is computed, and against which reference lines, is stated per block as `ref:LINE`
// = /root/reference/src/geometry/tract_geom_proc.py:LINE.  Nothing here is translated from the
// reference: the reference evaluates ~1,190 numpy calls per polyline and recomputes every shared
// intermediate 3-4 times; this file streams each polyline ONCE through a 4-stage register
// pipeline (point -> segment -> velocity -> binormal -> binormal derivative) and keeps every
// running sum in registers.
//
// Roofline note: B200 issues 64 fp64 lanes / SM / clock.  At 25.45 algorithmic bytes per point the
// HBM roof allows ~64 fp64 instructions per point; the metric set needs ~130.  The fp64 pipe, not
// HBM, bounds this kernel, so the design rule is "fewest fp64 instructions per point":
//   * one lane walks one polyline (or one contiguous chunk of a long one) sequentially — no
//     shuffles, no redundant halo work, no per-point reductions;
//   * sqrt / reciprocal are MUFU seeds + one cubic Newton step (full double precision, no IEEE
//     slow path), 1/(|d|+1e-12) is folded into the rsqrt;
//   * acos is a branch-light minimax polynomial on (1-c)/2.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tg {

constexpr double kEps = 1e-12;     // the additive epsilons of ref:58,68,77,80,92,102,140,145
constexpr double kMinLen = 1e-8;   // ref:40,45,160

// ------------------------------------------------------------------------------------------
// Elementary functions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double mufu_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double mufu_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
// 1/sqrt(x), x normal and positive.  Seed has ~20 good bits; one cubic step gives ~2^-58.
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y = mufu_rsqrt(x);
    double t = x * y;
    double e = fma(-t, y, 1.0);
    double p = fma(0.375, e, 0.5);
    double ye = y * e;
    return fma(ye, p, y);
}
// 1/x, x normal.  One cubic step: y(1 + e + e^2).
__device__ __forceinline__ double rcp_fast(double x) {
    double y = mufu_rcp(x);
    double e = fma(-x, y, 1.0);
    double e2 = fma(e, e, e);
    return fma(y, e2, y);
}
// true iff lo <= x < hi for positive normal doubles, decided on the high word (integer pipe).
__device__ __forceinline__ bool hi_in_range(double x, int lo_hi, int hi_hi) {
    return (unsigned)(__double2hiint(x) - lo_hi) < (unsigned)(hi_hi - lo_hi);
}
constexpr int kHi_1em6 = 0x3EB0C6F7;   // high word of 1e-6 (rounded down)
constexpr int kHi_1e300 = 0x7E37E43C;  // high word of 1e300
constexpr int kHi_1em280 = 0x05E0F3C5; // high word of ~1e-280
// len = sqrt(x2), inv = 1/(len + 1e-12)   (ref:102,145: d / (norm + 1e-12);  ref:58: norm + 1e-12)
__device__ __forceinline__ void norm_and_inv_eps(double x2, double& len, double& inv) {
    if (hi_in_range(x2, kHi_1em6 + 1, kHi_1e300)) {
        double r = rsqrt_fast(x2);
        len = x2 * r;
        double u = kEps * r;          // 1/(len+eps) = r/(1+eps r) = r - eps r^2 + O((eps r)^2 r), eps r < 1e-9
        inv = fma(-u, r, r);
    } else {
        len = sqrt(x2);
        inv = 1.0 / (len + kEps);
    }
}
__device__ __forceinline__ double sqrt_fast(double x) {   // x >= 0
    if (hi_in_range(x, kHi_1em280, kHi_1e300)) return x * rsqrt_fast(x);
    return sqrt(x);
}

// asin(s) = s + s*z*R(z), z = s*s in [0, 0.25]; degree-10 interpolant at Chebyshev nodes,
// |error| < 9e-16 relative on the resulting angle.
__device__ __forceinline__ double asin_core(double z) {
    double r = 0.027871289137110143;
    r = fma(r, z, -0.006822043980671263);
    r = fma(r, z, 0.015445133336819308);
    r = fma(r, z, 0.0102896411236249);
    r = fma(r, z, 0.014140941807431192);
    r = fma(r, z, 0.017337192543712077);
    r = fma(r, z, 0.022373010066676288);
    r = fma(r, z, 0.030381917485400308);
    r = fma(r, z, 0.044642857578717755);
    r = fma(r, z, 0.07499999999726302);
    r = fma(r, z, 0.1666666666666695);
    return r;
}
// arccos of c already clipped to [-1, 1] (ref:104-105).  NaN in -> NaN out.
__device__ __forceinline__ double acos_clipped(double c) {
    if (c > 0.5) {                       // the common case: consecutive segments nearly parallel
        double z = fma(-0.5, c, 0.5);    // (1-c)/2, exact for c in [0.5,1]
        double s = z * rsqrt_fast(z + 1e-300);   // z == 0 -> 0 * 1e150 = 0
        double zs = z * s;
        double t = fma(zs, asin_core(z), s);
        return t + t;
    } else if (c >= -0.5) {
        double z = c * c;
        double t = fma(c * z, asin_core(z), c);
        // pi/2 split hi/lo so that the subtraction is accurate
        return (1.5707963267948966 - t) + 6.123233995736766e-17;
    } else if (c >= -1.0) {
        double z = fma(0.5, c, 0.5);
        double s = z * rsqrt_fast(z + 1e-300);
        double t = fma(z * s, asin_core(z), s);
        return fma(-2.0, t, 3.141592653589793) + 1.2246467991473532e-16;
    }
    return c + c;  // NaN
}

// Eigenvalues of a symmetric 3x3 (a00 a01 a02 / a11 a12 / a22), descending (ref:122-123).
// Cyclic Jacobi: no trigonometry, relative accuracy on the small eigenvalues at least as good
// as LAPACK's (the trigonometric closed form fails the 1e-9 contract).
__device__ __forceinline__ void jacobi_rotate(double& app, double& aqq, double& apq, double& arp, double& arq) {
    // skip when |apq|^2 <= 1e-36 |app aqq|  (already diagonal to working precision) or apq is (sub)zero
    double thr = 1e-36 * fabs(app * aqq);
    double apq2 = apq * apq;
    if (!(apq2 > thr) || !(apq2 > 1e-290)) return;
    double d = aqq - app;
    double two = apq + apq;
    double h = fma(two, two, d * d);
    double den = fabs(d) + sqrt_fast(h);
    double t = copysign(two, (d >= 0.0) ? two : -two) * rcp_fast(den);   // tan of the rotation angle, |t| <= 1
    double c = rsqrt_fast(fma(t, t, 1.0));
    double s = t * c;
    double tapq = t * apq;
    app -= tapq;
    aqq += tapq;
    apq = 0.0;
    double rp = arp, rq = arq;
    arp = fma(c, rp, -(s * rq));
    arq = fma(s, rp, c * rq);
}
__device__ __forceinline__ void sym3_eigenvalues(double a00, double a01, double a02, double a11, double a12,
                                                 double a22, double& l1, double& l2, double& l3) {
#pragma unroll 1
    for (int sweep = 0; sweep < 8; ++sweep) {
        double off = a01 * a01 + a02 * a02 + a12 * a12;
        double dg = a00 * a00 + a11 * a11 + a22 * a22;
        if (!(off > 1e-26 * dg)) break;      // |a_pq| < 1e-13 |a|: the eigenvalue error left is ~|a_pq|^2 / gap
        jacobi_rotate(a00, a11, a01, a02, a12);   // (p,q)=(0,1), r=2: arp=a02, arq=a12
        jacobi_rotate(a00, a22, a02, a01, a12);   // (0,2), r=1: arp=a01, arq=a21
        jacobi_rotate(a11, a22, a12, a01, a02);   // (1,2), r=0: arp=a10, arq=a20
    }
    double hi = fmax(a00, a11), lo = fmin(a00, a11);
    l1 = fmax(hi, a22);
    l3 = fmin(lo, a22);
    l2 = fmax(lo, fmin(hi, a22));
}

// Non-iterative eigenvalues of a symmetric POSITIVE SEMI-DEFINITE 3x3 whose largest eigenvalue is isolated —
// the covariance of an elongated point cloud, i.e. nearly every polyline (ref:119-124).  Returns false
// (outputs untouched) when the matrix is not in that class; the caller then runs the Jacobi sweeps.
//   1. scale by 1/trace (eigenvalues now sum to 1);
//   2. lambda1 estimate from the trigonometric closed form in fp32 (accurate to ~1e-7: not good enough
//      as a result, SURVEY.md F5, but a fine Newton start), then two fp64 Newton steps on
//      det(A - x I) = 0 — the largest root is well conditioned when (l1 - l2) >= 0.05;
//   3. deflation: v1 = the largest cross product of two rows of A - l1 I, (u, w) an orthonormal basis
//      of its complement, and the 2x2 matrix [u w]^T A [u w] carries l2, l3 (closed form).
// Every step has absolute error O(eps * l1), the level of LAPACK's own (measured against 50-digit
// eigenvalues on 3000 random-walk polylines: ratios within 3.5e-12, LAPACK 2.5e-11; tests/test_highprec.py).
__device__ __forceinline__ bool sym3_eigenvalues_fast(const double c00, const double c01, const double c02,
                                                      const double c11, const double c12, const double c22,
                                                      double& l1, double& l2, double& l3) {
    const double tr = (c00 + c11) + c22;
    if (!hi_in_range(tr, kHi_1em280, kHi_1e300)) return false;       // also rejects NaN / negative / zero
    const double sc = rcp_fast(tr);
    const double a00 = c00 * sc, a01 = c01 * sc, a02 = c02 * sc, a11 = c11 * sc, a12 = c12 * sc, a22 = c22 * sc;
    // ---- fp32: closed form with the trace shifted out (q = 1/3)
    const float third = 0.33333334f;
    const float b00 = (float)a00 - third, b11 = (float)a11 - third, b22 = (float)a22 - third;
    const float f01 = (float)a01, f02 = (float)a02, f12 = (float)a12;
    const float off2 = fmaf(f01, f01, fmaf(f02, f02, f12 * f12));
    const float p2 = fmaf(2.0f, off2, fmaf(b00, b00, fmaf(b11, b11, b22 * b22)));
    if (!(p2 > 6e-6f)) return false;                                 // nearly isotropic: no isolated eigenvalue
    const float ip = rsqrtf(p2 * 0.16666667f);                       // 1/p, p = sqrt(p2/6)
    const float B00 = b00 * ip, B01 = f01 * ip, B02 = f02 * ip, B11 = b11 * ip, B12 = f12 * ip, B22 = b22 * ip;
    const float detB = fmaf(B00, fmaf(B11, B22, -(B12 * B12)), fmaf(-B01, fmaf(B01, B22, -(B12 * B02)), B02 * fmaf(B01, B12, -(B11 * B02))));
    const float r = fminf(fmaxf(0.5f * detB, -1.0f), 1.0f);
    const float phi = acosf(r) * 0.33333334f;
    const float two_p = 2.0f * __frcp_rn(ip);
    const float e1 = fmaf(two_p, __cosf(phi), third);
    const float e3 = fmaf(two_p, __cosf(phi + 2.0943951f), third);
    const float e2 = (1.0f - e1) - e3;                               // (e3 itself is not used as a result)
    if (!(e1 - e2 >= 0.05f)) return false;                           // l1 not isolated (e3 is only good to ~1e-7: the flatness test waits for fp64)
    // ---- fp64: Newton on x^3 - x^2 + k1 x - k0 (trace = 1)
    const double m0 = fma(a11, a22, -(a12 * a12)), m1 = fma(a00, a22, -(a02 * a02)), m2 = fma(a00, a11, -(a01 * a01));
    const double k1 = (m0 + m1) + m2;
    const double k0 = fma(a00, m0, fma(-a01, fma(a01, a22, -(a12 * a02)), a02 * fma(a01, a12, -(a11 * a02))));
    double x = (double)e1;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double px = fma(fma(x - 1.0, x, k1), x, -k0);
        const double dp = fma(fma(3.0, x, -2.0), x, k1);             // (x-l2)(x-l3)-ish: >= 0.05 * 0.05 near l1
        x = fma(-px, rcp_fast(dp), x);
    }
    // ---- eigenvector of x: largest cross product of two rows of A - x I
    const double d0 = a00 - x, d1 = a11 - x, d2 = a22 - x;
    const double p0x = fma(a01, a12, -(a02 * d1)), p0y = fma(a02, a01, -(d0 * a12)), p0z = fma(d0, d1, -(a01 * a01));    // r0 x r1
    const double p1x = fma(a01, d2, -(a02 * a12)), p1y = fma(a02, a02, -(d0 * d2)), p1z = fma(d0, a12, -(a01 * a02));    // r0 x r2
    const double p2x = fma(d1, d2, -(a12 * a12)), p2y = fma(a12, a02, -(a01 * d2)), p2z = fma(a01, a12, -(d1 * a02));    // r1 x r2
    const double n0 = fma(p0z, p0z, fma(p0y, p0y, p0x * p0x));
    const double n1 = fma(p1z, p1z, fma(p1y, p1y, p1x * p1x));
    const double n2 = fma(p2z, p2z, fma(p2y, p2y, p2x * p2x));
    double vx = p0x, vy = p0y, vz = p0z, nn = n0;
    if (n1 > nn) { vx = p1x; vy = p1y; vz = p1z; nn = n1; }
    if (n2 > nn) { vx = p2x; vy = p2y; vz = p2z; nn = n2; }
    if (!hi_in_range(nn, kHi_1em280, kHi_1e300)) return false;
    const double rn = rsqrt_fast(nn);
    vx *= rn; vy *= rn; vz *= rn;
    // ---- u: v x e_i normalised, i = the smallest |v_i| (so |u| before normalisation >= sqrt(2/3)); w = v x u
    const unsigned hx = (unsigned)__double2hiint(vx) & 0x7fffffffu, hy = (unsigned)__double2hiint(vy) & 0x7fffffffu,
                   hz = (unsigned)__double2hiint(vz) & 0x7fffffffu;
    double ux, uy, uz;
    if (hx <= hy && hx <= hz) { ux = 0.0; uy = -vz; uz = vy; }
    else if (hy <= hz) { ux = vz; uy = 0.0; uz = -vx; }
    else { ux = -vy; uy = vx; uz = 0.0; }
    const double ru = rsqrt_fast(fma(uz, uz, fma(uy, uy, ux * ux)));
    ux *= ru; uy *= ru; uz *= ru;
    const double wx = fma(vy, uz, -(vz * uy)), wy = fma(vz, ux, -(vx * uz)), wz = fma(vx, uy, -(vy * ux));
    // ---- 2x2 block of A in the (u, w) basis
    const double Aux = fma(a02, uz, fma(a01, uy, a00 * ux)), Auy = fma(a12, uz, fma(a11, uy, a01 * ux)), Auz = fma(a22, uz, fma(a12, uy, a02 * ux));
    const double Awx = fma(a02, wz, fma(a01, wy, a00 * wx)), Awy = fma(a12, wz, fma(a11, wy, a01 * wx)), Awz = fma(a22, wz, fma(a12, wy, a02 * wx));
    const double g00 = fma(uz, Auz, fma(uy, Auy, ux * Aux));
    const double g01 = fma(uz, Awz, fma(uy, Awy, ux * Awx));
    const double g11 = fma(wz, Awz, fma(wy, Awy, wx * Awx));
    const double h = 0.5 * (g00 + g11), d = 0.5 * (g00 - g11);
    const double rad2 = fma(d, d, g01 * g01);
    const double rad = hi_in_range(rad2, kHi_1em280, kHi_1e300) ? rad2 * rsqrt_fast(rad2) : sqrt(rad2);
    const double y2 = h + rad, y3 = h - rad;
    // absolute error ~5 eps l1 on every eigenvalue: within 1e-9 relative up to l1/l3 = 1e5 and far inside SURVEY.md N7's
    // |d lambda| <= 1e-12 l1 beyond; below l3 = 1e-9 l1 (planar or straight polylines, where the reference's `<= 1e-12 -> inf`
    // decisions hang on the small eigenvalues' last bits) the Jacobi sweeps, which are relatively accurate, take over
    if (!(y3 >= 1e-9) || !(x - y2 >= 0.04)) return false;
    l1 = x * tr; l2 = y2 * tr; l3 = y3 * tr;
    return true;
}

// ------------------------------------------------------------------------------------------
// Running sums of one polyline (or of one chunk of it).  Everything is additive across chunks
// except the curvature moments, which merge with Chan's formula.
// ------------------------------------------------------------------------------------------
struct Acc {
    double L;                 // sum |d_j|                                   ref:31-33
    double th;                // sum acos(clip(t_i . t_{i+1}))               ref:98-106
    double w0, w1, w2, ww;    // sum (t_j - r), sum |t_j - r|^2              ref:143-148 (shifted by r)
    double q0, q1, q2;        // sum (p_j - pm)                              ref:111-112, 119-121 (shifted by pm)
    double q00, q01, q02, q11, q12, q22;   // sum (p_j - pm)(p_j - pm)^T
    double mn0, mn1, mn2, mx0, mx1, mx2;   // ref:114-117
    double kK, k1, k2;        // curvature: shift, sum (k-K), sum (k-K)^2    ref:53-71
    double en;                // sum k_j^2 (|d_j| + 1e-12)                   ref:73-83
    double ta;                // sum of finite tau_j                          ref:85-96
    unsigned kn, tn;          // counts of finite kappa / tau
    unsigned absmax_hi;       // max over coordinates of (high word & 0x7fffffff): >= 0x7ff00000 <=> non-finite (ref:21)
};

__device__ __forceinline__ void acc_init(Acc& A) {
    A.L = A.th = A.w0 = A.w1 = A.w2 = A.ww = 0.0;
    A.q0 = A.q1 = A.q2 = A.q00 = A.q01 = A.q02 = A.q11 = A.q12 = A.q22 = 0.0;
    A.mn0 = A.mn1 = A.mn2 = __longlong_as_double(0x7ff0000000000000LL);
    A.mx0 = A.mx1 = A.mx2 = __longlong_as_double(0xfff0000000000000LL);
    A.kK = A.k1 = A.k2 = A.en = A.ta = 0.0;
    A.kn = A.tn = 0u;
    A.absmax_hi = 0u;
}

template <typename T>
__device__ __forceinline__ void load_point(const T* __restrict__ p, double& x, double& y, double& z) {
    x = (double)__ldg(p);
    y = (double)__ldg(p + 1);
    z = (double)__ldg(p + 2);
}

__device__ __forceinline__ bool finite_d(double x) {
    return (unsigned)(__double2hiint(x) & 0x7fffffff) < 0x7ff00000u;
}

// Stream points [max(c0-3,0) .. min(c1+2,n+2)] of a polyline of n >= 3 points through the
// pipeline and accumulate every quantity whose index lies in [c0, c1).
//   step k:  point k | segment k-1, angle k-2 | velocity k-1 | binormal, curvature k-2 | torsion k-3
// Edge rule (np.gradient, unit spacing, edge_order=1):
//   g_0 = f_1 - f_0,  g_j = (f_{j+1} - f_{j-1})/2,  g_{n-1} = f_{n-1} - f_{n-2}
// is realised as  g_j = s_j * (F(j+1) - F(j-1))  with F clamped to [0, n-1] and s_j = 1 at the two
// ends, 1/2 inside; multiplying by 1/2 is exact, so this is bit-identical to numpy's divide by 2.
template <typename T, bool WHOLE>
__device__ __forceinline__ void stream_chunk(const T* __restrict__ base, const int n, const int c0, const int c1,
                                             const double r0, const double r1, const double r2,
                                             const double m0, const double m1, const double m2, Acc& A) {
    const int last = n - 1;
    const int ks = WHOLE ? 0 : max(c0 - 3, 0);
    const int ke = WHOLE ? n + 2 : min(c1 + 2, n + 2);
    const bool torsion_on = n >= 4;                       // ref:86

    double cx, cy, cz;                                    // P(k)
    load_point(base + 3 * (int64_t)ks, cx, cy, cz);
    double p1x = cx, p1y = cy, p1z = cz;                  // P(k-1)
    double p2x = cx, p2y = cy, p2z = cz;                  // P(k-2)
    double tx = 0.0, ty = 0.0, tz = 0.0;                  // unit segment k-2
    double len_prev = 0.0;                                // |d_{k-2}|
    double vax = 0.0, vay = 0.0, vaz = 0.0;               // v_{k-2}
    double vbx = 0.0, vby = 0.0, vbz = 0.0;               // v_{k-3}
    double bax = 0.0, bay = 0.0, baz = 0.0, bba = 0.0;    // b_{k-3}, |b_{k-3}|^2
    double bbx = 0.0, bby = 0.0, bbz = 0.0;               // b_{k-4}

    double nx = cx, ny = cy, nz = cz;                     // P(k+1), prefetched one step ahead
    if (ks + 1 <= last) load_point(base + 3 * (int64_t)(ks + 1), nx, ny, nz);

#pragma unroll 1
    for (int k = ks; k <= ke; ++k) {
        // ---- prefetch P(k+2) while P(k) is being consumed
        double fx = nx, fy = ny, fz = nz;
        if (k + 2 <= last) load_point(base + 3 * (int64_t)(k + 2), fx, fy, fz);

        // ---- point stage, index k
        if (k <= last && (WHOLE || (k >= c0 && k < c1))) {
            unsigned hx = (unsigned)__double2hiint(cx) & 0x7fffffffu;
            unsigned hy = (unsigned)__double2hiint(cy) & 0x7fffffffu;
            unsigned hz = (unsigned)__double2hiint(cz) & 0x7fffffffu;
            A.absmax_hi = max(A.absmax_hi, max(hx, max(hy, hz)));
            A.mn0 = fmin(A.mn0, cx); A.mx0 = fmax(A.mx0, cx);
            A.mn1 = fmin(A.mn1, cy); A.mx1 = fmax(A.mx1, cy);
            A.mn2 = fmin(A.mn2, cz); A.mx2 = fmax(A.mx2, cz);
            double qx = cx - m0, qy = cy - m1, qz = cz - m2;
            A.q0 += qx; A.q1 += qy; A.q2 += qz;
            A.q00 = fma(qx, qx, A.q00); A.q01 = fma(qx, qy, A.q01); A.q02 = fma(qx, qz, A.q02);
            A.q11 = fma(qy, qy, A.q11); A.q12 = fma(qy, qz, A.q12); A.q22 = fma(qz, qz, A.q22);
        }

        // ---- segment stage, index j = k-1; bending angle i = k-2
        double len_cur = 0.0;
        if (k >= 1 && k <= last) {
            double dx = cx - p1x, dy = cy - p1y, dz = cz - p1z;          // ref:32
            double x2 = fma(dz, dz, fma(dy, dy, dx * dx));
            double inv;
            norm_and_inv_eps(x2, len_cur, inv);
            double ux = dx * inv, uy = dy * inv, uz = dz * inv;          // ref:102,145
            if (WHOLE || (k - 1 >= c0 && k - 1 < c1)) {
                A.L += len_cur;
                double wx = ux - r0, wy = uy - r1, wz = uz - r2;
                A.w0 += wx; A.w1 += wy; A.w2 += wz;
                A.ww = fma(wx, wx, fma(wy, wy, fma(wz, wz, A.ww)));
            }
            if (k >= 2 && (WHOLE || (k - 2 >= c0 && k - 2 < c1))) {
                double c = fma(tz, uz, fma(ty, uy, tx * ux));           // ref:103
                c = fmin(fmax(c, -1.0), 1.0);                           // ref:104
                A.th += acos_clipped(c);                                // ref:105-106 (acos >= 0, so |.| is a no-op)
            }
            tx = ux; ty = uy; tz = uz;
        }

        // ---- velocity stage, index j = k-1:  v_j = s_j (P(j+1) - P(j-1))        ref:49
        double vnx = vax, vny = vay, vnz = vaz;                          // hold v_{n-1} once j >= n
        if (k <= n) {
            double s = (k - 1 == 0 || k - 1 == last) ? 1.0 : 0.5;
            vnx = s * (cx - p2x); vny = s * (cy - p2y); vnz = s * (cz - p2z);
        }

        // ---- acceleration / binormal / curvature stage, index j = k-2            ref:50,57-59
        double bnx = bax, bny = bay, bnz = baz, bbn = bba;               // hold b_{n-1} once j >= n
        if (k >= 2 && k <= n + 1) {
            const int j = k - 2;
            double s = (j == 0 || j == last) ? 1.0 : 0.5;
            double ax = s * (vnx - vbx), ay = s * (vny - vby), az = s * (vnz - vbz);
            bnx = fma(vay, az, -(vaz * ay));                             // numpy cross, SURVEY.md N2
            bny = fma(vaz, ax, -(vax * az));
            bnz = fma(vax, ay, -(vay * ax));
            bbn = fma(bnz, bnz, fma(bny, bny, bnx * bnx));
            if (WHOLE || (j >= c0 && j < c1)) {
                double vv = fma(vaz, vaz, fma(vay, vay, vax * vax));
                double vlen, vinv;
                norm_and_inv_eps(vv, vlen, vinv);                        // vinv = 1/(|v| + 1e-12), ref:58
                double bmag = sqrt_fast(bbn);
                double kappa = bmag * (vinv * vinv * vinv);              // ref:59
                double kz = kappa;
                if (finite_d(kappa)) {                                   // ref:60,70
                    if (A.kn == 0u) A.kK = kappa;
                    double dk = kappa - A.kK;
                    A.k1 += dk;
                    A.k2 = fma(dk, dk, A.k2);
                    A.kn += 1u;
                } else {                                                 // ref:81 nan_to_num
                    kz = (kappa != kappa) ? 0.0 : copysign(1.7976931348623157e308, kappa);
                }
                if (j < last) A.en = fma(kz * kz, len_prev + kEps, A.en);   // ref:77,82-83: ds_j = |d_j| + 1e-12
            }
        }

        // ---- binormal-derivative / torsion stage, index j = k-3                   ref:91-95
        if (torsion_on && k >= 3 && (WHOLE || (k - 3 >= c0 && k - 3 < c1))) {
            const int j = k - 3;
            double s = (j == 0 || j == last) ? 1.0 : 0.5;
            double ex = s * (bnx - bbx), ey = s * (bny - bby), ez = s * (bnz - bbz);
            double num = fma(baz, ez, fma(bay, ey, bax * ex));           // ref:93
            double den = bba + kEps;                                     // ref:92
            double tau = hi_in_range(den, kHi_1em280, kHi_1e300) ? num * rcp_fast(den) : num / den;
            if (finite_d(tau)) { A.ta += tau; A.tn += 1u; }              // ref:95
        }

        // ---- rotate the pipeline registers (left clamp: F(-1) := F(0))
        if (k == 1) { vbx = vnx; vby = vny; vbz = vnz; } else { vbx = vax; vby = vay; vbz = vaz; }
        vax = vnx; vay = vny; vaz = vnz;
        if (k == 2) { bbx = bnx; bby = bny; bbz = bnz; } else { bbx = bax; bby = bay; bbz = baz; }
        bax = bnx; bay = bny; baz = bnz; bba = bbn;
        len_prev = len_cur;
        p2x = p1x; p2y = p1y; p2z = p1z;
        p1x = cx; p1y = cy; p1z = cz;
        cx = nx; cy = ny; cz = nz;                                        // right clamp: P(k) := P(n-1) for k >= n
        nx = fx; ny = fy; nz = fz;
    }
}

// Merge chunk partial B into A (A covers lower indices).  Chan et al. for the curvature moments.
__device__ __forceinline__ void acc_merge(Acc& A, const Acc& B) {
    A.L += B.L; A.th += B.th;
    A.w0 += B.w0; A.w1 += B.w1; A.w2 += B.w2; A.ww += B.ww;
    A.q0 += B.q0; A.q1 += B.q1; A.q2 += B.q2;
    A.q00 += B.q00; A.q01 += B.q01; A.q02 += B.q02; A.q11 += B.q11; A.q12 += B.q12; A.q22 += B.q22;
    A.mn0 = fmin(A.mn0, B.mn0); A.mn1 = fmin(A.mn1, B.mn1); A.mn2 = fmin(A.mn2, B.mn2);
    A.mx0 = fmax(A.mx0, B.mx0); A.mx1 = fmax(A.mx1, B.mx1); A.mx2 = fmax(A.mx2, B.mx2);
    A.en += B.en; A.ta += B.ta; A.tn += B.tn;
    A.absmax_hi = max(A.absmax_hi, B.absmax_hi);
    if (B.kn != 0u) {
        if (A.kn == 0u) { A.kK = B.kK; A.k1 = B.k1; A.k2 = B.k2; A.kn = B.kn; }
        else {
            // re-express B's shifted sums around A's shift: k - KA = (k - KB) + (KB - KA)
            double dK = B.kK - A.kK;
            double nb = (double)B.kn;
            A.k2 += B.k2 + dK * (2.0 * B.k1 + nb * dK);
            A.k1 += B.k1 + nb * dK;
            A.kn += B.kn;
        }
    }
}

// Turn the running sums of a complete polyline (n >= 3) into the 17 metrics (column-major out).
// Returns the keep flags.  p_first / p_last are the end points (ref:36).
__device__ __forceinline__ unsigned finalize_metrics(const Acc& A, const int n, const double f0, const double f1,
                                                     const double f2, const double e0, const double e1, const double e2,
                                                     const double m0, const double m1, const double m2,
                                                     double* __restrict__ out, const int64_t S, const int64_t s) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    unsigned keep = 0u;
    if (A.absmax_hi < 0x7ff00000u) keep |= 1u;            // ref:21 (n > 2 checked by the caller)
    if (A.L > kMinLen) keep |= 2u;                        // ref:160
    if (keep != 3u) {
#pragma unroll
        for (int m = 0; m < 17; ++m) out[(int64_t)m * S + s] = nan;
        return keep;
    }
    const double dn = (double)n;
    const double L = A.L;
    double cx = e0 - f0, cy = e1 - f1, cz = e2 - f2;
    double chord = sqrt(cx * cx + cy * cy + cz * cz);                      // ref:36
    out[0 * S + s] = L;
    out[1 * S + s] = chord;
    out[2 * S + s] = L / fmax(chord, kMinLen);                             // ref:38-41
    out[3 * S + s] = chord / fmax(L, kMinLen);                             // ref:43-46
    double kmean = 0.0, kstd = 0.0;
    if (A.kn != 0u) {                                                      // ref:61,71
        double kn = (double)A.kn;
        double dm = A.k1 / kn;
        kmean = A.kK + dm;
        double m2c = A.k2 - A.k1 * dm;
        kstd = sqrt(fmax(m2c, 0.0) / kn);
    }
    out[4 * S + s] = kmean;
    out[5 * S + s] = kstd;
    out[6 * S + s] = A.en;
    out[7 * S + s] = (n >= 4 && A.tn != 0u) ? A.ta / (double)A.tn : 0.0;   // ref:86-87,96
    out[8 * S + s] = A.th / (double)(n - 2);                               // ref:106
    out[9 * S + s] = ((A.mx0 - A.mn0) * (A.mx1 - A.mn1)) * (A.mx2 - A.mn2);   // ref:117
    // covariance, ddof=1, scaled by the reciprocal as numpy does
    double g0 = A.q0 / dn, g1 = A.q1 / dn, g2 = A.q2 / dn;                 // centroid - pm
    double f = 1.0 / (dn - 1.0);
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
    double dm1 = (double)(n - 1);
    double a0 = A.w0 / dm1, a1 = A.w1 / dm1, a2 = A.w2 / dm1;
    double disp = A.ww / dm1 - (a0 * a0 + a1 * a1 + a2 * a2);              // ref:146-147, mean|w|^2 - |mean w|^2
    out[16 * S + s] = fmax(disp, 0.0);
    return keep;
}

}  // namespace tg
