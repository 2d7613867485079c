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
template <> struct Num<double> {
    static D3D_HD double rcp(double x) { return 1.0 / x; }
    static D3D_HD double sat(double x) { return fmin(fmax(x, 0.0), 1.0); }  // NaN -> 0
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
    static D3D_HD double fmin_(double a, double b) { return fmin(a, b); }
    static D3D_HD double fmax_(double a, double b) { return fmax(a, b); }
    static D3D_HD double abs_(double a) { return fabs(a); }
};

// signed integral contribution of one directed edge P->Q of A (in B's frame):
//   (x0 - x1) * mean over the edge's clamped span of clamp(y, -hh, hh),  x0/x1 = clamp(Px/Qx, -hw, hw)
template <typename T>
D3D_HD T edge_term(T px, T py, T qx, T qy, T slope, T hw, T hh)
{
    using N = Num<T>;
    T x0 = N::fmin_(N::fmax_(px, -hw), hw);
    T x1 = N::fmin_(N::fmax_(qx, -hw), hw);
    T ya = py + slope * (x0 - px);
    T yb = qy + slope * (x1 - qx);
    T ymin = N::fmin_(ya, yb), ymax = N::fmax_(ya, yb);
    T inv = N::rcp(ymax - ymin);                     // +inf for a horizontal span: fractions become 0/1
    T flo = N::sat((-hh - ymin) * inv);              // fraction of the span below -hh  (sat(NaN) = 0)
    T fhi = N::sat((ymax - hh) * inv);               // fraction above +hh
    T cmin = N::fmin_(N::fmax_(ymin, -hh), hh);
    T cmax = N::fmin_(N::fmax_(ymax, -hh), hh);
    T mean = hh * (fhi - flo) + (T(1) - flo - fhi) * (T(0.5) * (cmin + cmax));
    return (x0 - x1) * mean;
}

// intersection-over-union of two rotated boxes given their records.  ~150 instructions, no branches.
template <typename T>
D3D_HD T rbox_iou(const BoxRec<T> &A, const BoxRec<T> &B)
{
    using N = Num<T>;
    // A's centre and axes in B's frame
    T dx = A.cx - B.cx, dy = A.cy - B.cy;
    T cx = B.c * dx + B.s * dy;
    T cy = B.c * dy - B.s * dx;
    T c = A.c * B.c + A.s * B.s;   // cos(rA - rB)
    // sin(rA - rB) from two separately rounded products: exactly 0 for equal headings (an FMA would leave
    // the rounding residue of one product, i.e. a 1e-8 rad tilt that costs 1e-4 of IoU on 1e4:1 slivers)
    T s = N::mul_rn(A.s, B.c) - N::mul_rn(A.c, B.s);
    T ux = c * A.hw, uy = s * A.hw;    // half width axis
    T vx = -s * A.hh, vy = c * A.hh;   // half height axis
    // CCW vertices V0 = C-u-v, V1 = C+u-v, V2 = C+u+v, V3 = C-u+v  (same order as geometry.hpp:417-429)
    T mx = cx - ux, px = cx + ux, my = cy - uy, py = cy + uy;
    T x0 = mx - vx, y0 = my - vy;
    T x1 = px - vx, y1 = py - vy;
    T x2 = px + vx, y2 = py + vy;
    T x3 = mx + vx, y3 = my + vy;
    // slopes dy/dx of the u-edges (s/c) and v-edges (-c/s); a vertical edge has zero clamped x-extent,
    // so any finite slope works for it
    T mu = N::abs_(c) > N::tiny() ? s * N::rcp(c) : T(0);
    T mv = N::abs_(s) > N::tiny() ? -c * N::rcp(s) : T(0);
    T acc = edge_term<T>(x0, y0, x1, y1, mu, B.hw, B.hh);
    acc += edge_term<T>(x1, y1, x2, y2, mv, B.hw, B.hh);
    acc += edge_term<T>(x2, y2, x3, y3, mu, B.hw, B.hh);
    acc += edge_term<T>(x3, y3, x0, y0, mv, B.hw, B.hh);
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
