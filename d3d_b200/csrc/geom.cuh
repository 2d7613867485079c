// geom.cuh -- per-box records and the register-resident rotated-rectangle intersection used by the
// IoU tile kernel and the NMS mask kernel.
//
// Algorithm (ours, not dgal's): move box A into box B's frame, where B is the axis-aligned rectangle
// [-hw,hw] x [-hh,hh].  With A's boundary traversed counter-clockwise,
//     area(A ^ B) = - sum over A's 4 edges  of  integral_{x in edge, clamped to [-hw,hw]} clamp(y_edge(x), -hh, hh) dx
// (the y-interval of A at abscissa x is [ylo(x), yhi(x)]; its overlap with [-hh,hh] has length
// clamp(yhi) - clamp(ylo); upper edges run right-to-left and lower edges left-to-right, so the signed
// edge integrals add up to the overlap).  Each edge integral is the edge's clamped x-extent times the
// mean of clamp(y) over the edge's y-range, which is a closed form of saturating fractions.  The sum
// is a continuous function of the inputs with no topological decisions (no vertex lists, no
// inside/outside classification), so parallel, touching and coincident edges -- the cases where the
// reference's Rotating-Calipers default returns 1.0 or garbage -- degrade
// gracefully instead of flipping; it needs no dynamically indexed storage, so everything stays in
// registers; and it is branch-free, so a warp of 32 candidate pairs never diverges.
// Reference behaviour replaced: dgal::poly2_from_xywhr + intersect + area + iou
// (thirdparty/dgal/geometry.hpp:417-429, 686-842, 921-932, 1215-1224).
#pragma once
#include "common.cuh"
#include <math.h>

namespace d3d {

// one record per box, computed once per call (the reference rebuilds the quad from xywhr for every pair)
template <typename T> struct __align__(16) BoxRec {
    T cx, cy;    // centre
    T c, s;      // cos r, sin r
    T hw, hh;    // half extents (|w|/2, |h|/2)
    T rho;       // bounding-circle radius; NaN marks a padding record (fails every candidate test)
    T area;      // w*h
};

// axis-aligned bounding box of the rotated quad, for method="box" (geometry.hpp:398-414)
template <typename T> struct __align__(16) AABBRec { T minx, maxx, miny, maxy; };

// D3D_HD: the clip math is also compilable as host code so that the CPU test-suite can check the
// ALGORITHM on machines without a GPU.  The product never runs it on the host.
#define D3D_HD __host__ __device__ __forceinline__

template <typename T> struct Num;
template <> struct Num<float> {
    static D3D_HD float rcp(float x)
    {
#ifdef __CUDA_ARCH__
        float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;   // MUFU.RCP, ~1 ulp
#else
        return 1.0f / x;
#endif
    }
    static D3D_HD float sat(float x)
    {
#ifdef __CUDA_ARCH__
        return __saturatef(x);                     // NaN -> +0
#else
        return fminf(fmaxf(x, 0.0f), 1.0f);        // C99 fmaxf(NaN, 0) = 0
#endif
    }
    static D3D_HD float mul_rn(float a, float b)   // product that is never contracted into an FMA
    {
#ifdef __CUDA_ARCH__
        return __fmul_rn(a, b);
#else
        return a * b;
#endif
    }
    static D3D_HD float tiny() { return 1e-30f; }
    static D3D_HD float snap() { return 1e-6f; }
    static D3D_HD float fmin_(float a, float b) { return fminf(a, b); }
    static D3D_HD float fmax_(float a, float b) { return fmaxf(a, b); }
    static D3D_HD float abs_(float a) { return fabsf(a); }
};
// Double precision on the device: fmin / fmax cost 7 instructions each on sm_100a (DSETP.MIN + selects + NaN canonicalisation, there
// is no DMNMX) and 1.0 / x is MUFU.RCP64H + five DFMA behind a range check with a call for zero, infinite and subnormal arguments --
// a horizontal edge (d = 0) of an axis-aligned box takes that call.  The clip only needs: min / max of finite numbers, sat with
// NaN -> 0, and a reciprocal that is +-inf for 0 and good to an ulp elsewhere; written as compare + select (3 instructions) and as
// RCP64H + two Newton steps without a branch they take a fifth off the clip's instruction count.
template <> struct Num<double> {
    static D3D_HD double rcp(double x)
    {
#ifdef __CUDA_ARCH__
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));   // ~20 bits; 0 (and subnormals) -> inf, inf -> 0
        const double e0 = fma(-x, r, 1.0), q0 = fma(r, e0, r);
        const double e1 = fma(-x, q0, 1.0), q1 = fma(q0, e1, q0);
        const double ar = fabs(r);
        return (ar > 0.0 && ar < 1.7976931348623157e308) ? q1 : r;   // the Newton steps turn 0 and inf into NaN: keep the seed there
#else
        return 1.0 / x;
#endif
    }
    static D3D_HD double sat(double x)   // NaN -> 0
    {
#ifdef __CUDA_ARCH__
        const double lo = x > 0.0 ? x : 0.0;
        return lo < 1.0 ? lo : 1.0;
#else
        return fmin(fmax(x, 0.0), 1.0);
#endif
    }
    static D3D_HD double mul_rn(double a, double b)
    {
#ifdef __CUDA_ARCH__
        return __dmul_rn(a, b);
#else
        return a * b;
#endif
    }
    static D3D_HD double tiny() { return 1e-280; }
    static D3D_HD double snap() { return 1e-12; }
#ifdef __CUDA_ARCH__
    static D3D_HD double fmin_(double a, double b) { return a < b ? a : b; }   // (finite arguments)
    static D3D_HD double fmax_(double a, double b) { return a > b ? a : b; }
#else
    static D3D_HD double fmin_(double a, double b) { return fmin(a, b); }
    static D3D_HD double fmax_(double a, double b) { return fmax(a, b); }
#endif
    static D3D_HD double abs_(double a) { return fabs(a); }
};

// Pair arithmetic.  On sm_100a a float2 add / mul / fma is ONE instruction (FADD2 / FMUL2 / FFMA2: two fp32 results per
// lane per issue slot, operand negation, half swaps and scalar broadcasts are free operand modifiers), and the clip
// below is issue-bound, so its edge integrals are evaluated two edges at a time.  For double (and on the host, where
// the CPU test-suite checks the algorithm) the same code is two scalar operations.
template <typename T> struct Vec2;
template <> struct Vec2<float>  { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename T> struct P2 {
    using V = typename Vec2<T>::type;
    static D3D_HD V mk(T x, T y) { V r; r.x = x; r.y = y; return r; }
    static D3D_HD V add(V a, V b)
    {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
        if constexpr (sizeof(T) == 4) return __fadd2_rn(a, b);
#endif
        return mk(a.x + b.x, a.y + b.y);
    }
    static D3D_HD V sub(V a, V b)
    {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
        if constexpr (sizeof(T) == 4) return __fadd2_rn(a, mk(-b.x, -b.y));
#endif
        return mk(a.x - b.x, a.y - b.y);
    }
    static D3D_HD V mul(V a, V b)
    {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
        if constexpr (sizeof(T) == 4) return __fmul2_rn(a, b);
#endif
        return mk(a.x * b.x, a.y * b.y);
    }
    static D3D_HD V fma(V a, V b, V c)     // a*b + c
    {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
        if constexpr (sizeof(T) == 4) return __ffma2_rn(a, b, c);
#endif
        return mk(a.x * b.x + c.x, a.y * b.y + c.y);
    }
    static D3D_HD V fnma(V a, V b, V c)    // c - a*b
    {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
        if constexpr (sizeof(T) == 4) return __ffma2_rn(mk(-a.x, -a.y), b, c);
#endif
        return mk(c.x - a.x * b.x, c.y - a.y * b.y);
    }
};

// Two directed edges of A at once (in B's frame).  ya / yb: the edge's ordinate at its start / end abscissa after the
// abscissae were clamped to [-hw, hw].  Returns TWICE the mean over each edge's clamped span of clamp(y, -hh, hh):
//   flo / fhi = fractions of the span below -hh / above +hh;  the part inside runs from cmin = ymin + flo*d to
//   cmax = ymax - fhi*d, so cmin + cmax = ya + yb - d*(fhi - flo);   2*mean = 2*hh*(fhi - flo) + (1 - flo - fhi)*(cmin + cmax)
template <typename T>
D3D_HD typename P2<T>::V span_mean2(typename P2<T>::V ya, typename P2<T>::V yb, T hh)
{
    using N = Num<T>; using P = P2<T>; using V = typename P::V;
    const V ymin = P::mk(N::fmin_(ya.x, yb.x), N::fmin_(ya.y, yb.y));
    const V ymax = P::mk(N::fmax_(ya.x, yb.x), N::fmax_(ya.y, yb.y));
    const V d = P::sub(ymax, ymin);
    const T ix = N::rcp(d.x), iy = N::rcp(d.y);          // +inf for a horizontal span: the fractions become 0 / 1
    const V nhh = P::mk(-hh, -hh);
    const V tl = P::sub(nhh, ymin), th = P::add(ymax, nhh);
    const V flo = P::mk(N::sat(tl.x * ix), N::sat(tl.y * iy));   // sat(NaN) = 0
    const V fhi = P::mk(N::sat(th.x * ix), N::sat(th.y * iy));
    const V diff = P::sub(fhi, flo);
    const V csum = P::fnma(d, diff, P::add(ya, yb));
    const V wmid = P::sub(P::sub(P::mk(T(1), T(1)), flo), fhi);
    return P::fma(P::mk(hh + hh, hh + hh), diff, P::mul(wmid, csum));
}

// intersection-over-union of two rotated boxes given their records.  No branches, registers only.
template <typename T>
D3D_HD T rbox_iou(const BoxRec<T> &A, const BoxRec<T> &B)
{
    using N = Num<T>; using P = P2<T>; using V = typename P::V;
    // A's centre and axes in B's frame
    T dx = A.cx - B.cx, dy = A.cy - B.cy;
    T cx = B.c * dx + B.s * dy;
    T cy = B.c * dy - B.s * dx;
    T c = A.c * B.c + A.s * B.s;   // cos(rA - rB)
    // sin(rA - rB) from two separately rounded products: exactly 0 for equal headings (an FMA would leave
    // the rounding residue of one product, i.e. a 1e-8 rad tilt that costs 1e-4 of IoU on 1e4:1 slivers)
    T s = N::mul_rn(A.s, B.c) - N::mul_rn(A.c, B.s);
    // half axes u (width) and v (height); a = u + v, b = u - v
    const V u = P::mul(P::mk(c, s), P::mk(A.hw, A.hw)), v = P::mul(P::mk(-s, c), P::mk(A.hh, A.hh));
    const V a = P::add(u, v), b = P::sub(u, v);
    // CCW vertices V0 = C-a, V1 = C+b, V2 = C+a, V3 = C-b (same order as geometry.hpp:417-429), held as the pairs
    // (V0, V2) and (V1, V3): edges e0 = V0->V1 and e2 = V2->V3 share the slope s/c, e1 = V1->V2 and e3 = V3->V0 share -c/s
    const V pm = P::mk(T(-1), T(1)), mp = P::mk(T(1), T(-1));
    const V X02 = P::fma(P::mk(a.x, a.x), pm, P::mk(cx, cx)), Y02 = P::fma(P::mk(a.y, a.y), pm, P::mk(cy, cy));
    const V X13 = P::fma(P::mk(b.x, b.x), mp, P::mk(cx, cx)), Y13 = P::fma(P::mk(b.y, b.y), mp, P::mk(cy, cy));
    // a vertical edge has zero clamped x-extent, so any finite slope works for it
    const T mu = N::abs_(c) > N::tiny() ? s * N::rcp(c) : T(0);
    const T mv = N::abs_(s) > N::tiny() ? -c * N::rcp(s) : T(0);
    const T hw = B.hw, hh = B.hh;
    const V CX02 = P::mk(N::fmin_(N::fmax_(X02.x, -hw), hw), N::fmin_(N::fmax_(X02.y, -hw), hw));
    const V CX13 = P::mk(N::fmin_(N::fmax_(X13.x, -hw), hw), N::fmin_(N::fmax_(X13.y, -hw), hw));
    const V D02 = P::sub(CX02, X02), D13 = P::sub(CX13, X13);       // how far the clamp moved each vertex abscissa
    const V mu2 = P::mk(mu, mu), mv2 = P::mk(mv, mv);
    const V ya02 = P::fma(mu2, D02, Y02), yb02 = P::fma(mu2, D13, Y13);   // e0, e2: start (V0, V2), end (V1, V3)
    const V ya13 = P::fma(mv2, D13, Y13), t = P::fma(mv2, D02, Y02);      // e1, e3: start (V1, V3), end (V2, V0)
    const V yb13 = P::mk(t.y, t.x);
    // signed area under the clamped boundary: sum over the edges of (x_start - x_end) * mean clamp(y)
    const V w02 = P::sub(CX02, CX13), w13 = P::sub(CX13, P::mk(CX02.y, CX02.x));
    const V acc2 = P::fma(w13, span_mean2<T>(ya13, yb13, hh), P::mul(w02, span_mean2<T>(ya02, yb02, hh)));
    const T acc = T(0.5) * (acc2.x + acc2.y);
    // the four edge integrals of two disjoint boxes cancel only up to rounding: snap residues below
    // snap()*(areaA+areaB) (IoU error <= 2*snap) to exactly +0 like the reference's empty intersection
    T asum = A.area + B.area;
    T ai = acc > N::snap() * asum ? acc : T(0);
    T au = asum - ai;
    // two zero-area boxes: the reference divides 0/0 (geometry.hpp:1223); we return 0
    return au > T(0) ? ai * N::rcp(au) : T(0);
}


// box row (x,y,w,h,r) -> record.  sincos of the unreduced heading is evaluated once per box.
template <typename T>
D3D_HD BoxRec<T> make_box_rec(T x, T y, T w, T h, T r)
{
    BoxRec<T> b;
    b.cx = x; b.cy = y;
    b.c = cos(r); b.s = sin(r);
    T aw = Num<T>::abs_(w), ah = Num<T>::abs_(h);
    b.hw = T(0.5) * aw; b.hh = T(0.5) * ah;
    b.rho = T(0.5) * sqrt(aw * aw + ah * ah);
    b.area = aw * ah;
    return b;
}

// AABB of the rotated quad with the reference's own vertex formulas (geometry.hpp:417-429) so that
// method="box" reproduces its extents to the last bit in fp64
template <typename T>
D3D_HD AABBRec<T> make_aabb_rec(T x, T y, T w, T h, T r)
{
    using N = Num<T>;
    T sr = sin(r), cr = cos(r);
    T dxsin = w * sr / 2, dxcos = w * cr / 2, dysin = h * sr / 2, dycos = h * cr / 2;
    T x0 = x - dxcos + dysin, y0 = y - dxsin - dycos;
    T x1 = x + dxcos + dysin, y1 = y + dxsin - dycos;
    T x2 = x + dxcos - dysin, y2 = y + dxsin + dycos;
    T x3 = x - dxcos - dysin, y3 = y - dxsin + dycos;
    AABBRec<T> a;
    a.minx = N::fmin_(N::fmin_(x0, x1), N::fmin_(x2, x3)); a.maxx = N::fmax_(N::fmax_(x0, x1), N::fmax_(x2, x3));
    a.miny = N::fmin_(N::fmin_(y0, y1), N::fmin_(y2, y3)); a.maxy = N::fmax_(N::fmax_(y0, y1), N::fmax_(y2, y3));
    return a;
}

// IoU of two AABBs: geometry.hpp:513-529 (empty when they only touch), :914-918, :1206-1212
template <typename T>
D3D_HD T aabb_iou(const AABBRec<T> &a, const AABBRec<T> &b)
{
    using N = Num<T>;
    T ai = T(0);
    if (!(a.maxx <= b.minx || a.minx >= b.maxx || a.maxy <= b.miny || a.miny >= b.maxy))
        ai = (N::fmin_(a.maxx, b.maxx) - N::fmax_(a.minx, b.minx)) * (N::fmin_(a.maxy, b.maxy) - N::fmax_(a.miny, b.miny));
    T au = (a.maxx - a.minx) * (a.maxy - a.miny) + (b.maxx - b.minx) * (b.maxy - b.miny) - ai;
    return ai / au;
}

}  // namespace d3d
