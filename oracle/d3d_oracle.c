/* d3d_oracle.c -- CPU oracle for the d3d hot path (rotated IoU, NMS, voxelization, aligned scatter).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C *restatement* of the reference's CPU algorithms,
 * written to be the checker the CUDA path is compared against.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; nothing under d3d_b200/ does.
 * Parity status: PINNED -- tests/test_oracle_pinned.py checks it against (a) the reference's own
 * golden vectors and known answers (test/voxel_data.npz, test/test_box.py, test/test_point.py values,
 * committed under tests/golden/), (b) fixtures generated from the reference's own compiled CPU
 * extensions (oracle/_ref, script tests/golden/make_golden.py) and (c) oracle/_ref live, when present.
 *
 * Build: `make -C oracle` -> oracle/libd3d_oracle.so   (gcc -O2 -ffp-contract=off: no FMA contraction,
 * so the double paths reproduce the reference's x86-64 -O2 build bit for bit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ polygon arithmetic, 3 precisions */
#define T float
#define SFX f
#define EPS 3e-7 /* Numeric<float>::eps(), geometry.hpp:103-108 */
#define M_SIN sinf
#define M_COS cosf
#define M_ATAN2 atan2f
#define M_HYPOT hypotf
#define M_FABS fabsf
#include "geom_body.inc"
#undef T
#undef SFX
#undef EPS
#undef M_SIN
#undef M_COS
#undef M_ATAN2
#undef M_HYPOT
#undef M_FABS

#define T double
#define SFX d
#define EPS 6e-15 /* Numeric<double>::eps(), geometry.hpp:109-114 */
#define M_SIN sin
#define M_COS cos
#define M_ATAN2 atan2
#define M_HYPOT hypot
#define M_FABS fabs
#include "geom_body.inc"
#undef T
#undef SFX
#undef EPS
#undef M_SIN
#undef M_COS
#undef M_ATAN2
#undef M_HYPOT
#undef M_FABS

#define T long double
#define SFX l
#define EPS 0.0L
#define M_SIN sinl
#define M_COS cosl
#define M_ATAN2 atan2l
#define M_HYPOT hypotl
#define M_FABS fabsl
#include "geom_body.inc"
#undef T
#undef SFX
#undef EPS
#undef M_SIN
#undef M_COS
#undef M_ATAN2
#undef M_HYPOT
#undef M_FABS

enum { ORC_ALG_RC = 1, ORC_ALG_SH = 2, ORC_ALG_TRUTH = 3 };
enum { ORC_IOU_BOX = 1, ORC_IOU_RBOX = 2 };           /* d3d/box/common.h:5-9 */
enum { ORC_SUP_HARD = 0, ORC_SUP_LINEAR = 1, ORC_SUP_GAUSSIAN = 2 }; /* d3d/box/common.h:10 */

/* ---- pairwise rotated IoU: d3d/box/iou.cpp:94-123 iou2dr_forward_templated.  out is [n,m] row-major.
 * blowups (optional, u8[n,m]) flags pairs where RC walked past dgal's buffers. */
void orc_iou2dr_f32(const float *b1, int64_t n, const float *b2, int64_t m, float *out, int alg, uint8_t *blowups)
{
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) {
            int bl = 0;
            out[i * m + j] = quad_iou_f(b1 + 5 * i, b2 + 5 * j, alg, &bl);
            if (blowups) blowups[i * m + j] = (uint8_t)bl;
        }
}
void orc_iou2dr_f64(const double *b1, int64_t n, const double *b2, int64_t m, double *out, int alg, uint8_t *blowups)
{
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) {
            int bl = 0;
            out[i * m + j] = quad_iou_d(b1 + 5 * i, b2 + 5 * j, alg, &bl);
            if (blowups) blowups[i * m + j] = (uint8_t)bl;
        }
}
/* geometric truth in long double from double inputs (the authority for degenerate pairs) */
void orc_iou2dr_truth(const double *b1, int64_t n, const double *b2, int64_t m, double *out)
{
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) {
            long double a[5], b[5];
            for (int k = 0; k < 5; k++) { a[k] = b1[5 * i + k]; b[k] = b2[5 * j + k]; }
            out[i * m + j] = (double)quad_iou_l(a, b, ORC_ALG_TRUTH, NULL);
        }
}
/* ---- pairwise AABB IoU: d3d/box/iou.cpp:11-46 iou2d_forward */
void orc_iou2d_f32(const float *b1, int64_t n, const float *b2, int64_t m, float *out)
{
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) out[i * m + j] = aabb_iou_f(b1 + 5 * i, b2 + 5 * j);
}
void orc_iou2d_f64(const double *b1, int64_t n, const double *b2, int64_t m, double *out)
{
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) out[i * m + j] = aabb_iou_d(b1 + 5 * i, b2 + 5 * j);
}

/* ---- point-in-rotated-box mask: d3d/box/utils.cpp:10-47 crop_2dr.  out is u8[m boxes, n points].
 * Vertices as dgal::poly2_from_xywhr (geometry.hpp:417-429), bounding box as aabox2_from_poly2 (:398-414) with the open
 * test of AABox2::contains (:185-188), then Poly2::contains (:218-229): no edge may have _cross(...) < 0 (:160-163). */
#define ORC_CROP_BODY(T, SFX, SIN, COS)                                                                       \
    void orc_crop2dr_##SFX(const T *pts, int64_t n, const T *boxes, int64_t m, uint8_t *out)                  \
    {                                                                                                         \
        for (int64_t i = 0; i < m; i++) {                                                                     \
            const T x = boxes[5 * i], y = boxes[5 * i + 1], w = boxes[5 * i + 2], h = boxes[5 * i + 3], r = boxes[5 * i + 4]; \
            const T dxsin = w * SIN(r) / 2, dxcos = w * COS(r) / 2, dysin = h * SIN(r) / 2, dycos = h * COS(r) / 2; \
            const T vx[4] = {x - dxcos + dysin, x + dxcos + dysin, x + dxcos - dysin, x - dxcos - dysin};       \
            const T vy[4] = {y - dxsin - dycos, y + dxsin - dycos, y + dxsin + dycos, y - dxsin + dycos};       \
            T minx = vx[0], maxx = vx[0], miny = vy[0], maxy = vy[0];                                         \
            for (int k = 1; k < 4; k++) {                                                                     \
                if (vx[k] < minx) minx = vx[k];                                                               \
                if (vx[k] > maxx) maxx = vx[k];                                                               \
                if (vy[k] < miny) miny = vy[k];                                                               \
                if (vy[k] > maxy) maxy = vy[k];                                                               \
            }                                                                                                 \
            for (int64_t j = 0; j < n; j++) {                                                                 \
                const T px = pts[2 * j], py = pts[2 * j + 1];                                                 \
                int in = px > minx && px < maxx && py > miny && py < maxy;                                    \
                for (int k = 0; k < 4 && in; k++) {                                                           \
                    const int a = (k + 3) & 3;   /* edge a -> k; order 3->0, 0->1, 1->2, 2->3 as geometry.hpp:221-226 */ \
                    const T c = (vx[k] - vx[a]) * (py - vy[k]) - (vy[k] - vy[a]) * (px - vx[k]);              \
                    if (c < 0) in = 0;                                                                        \
                }                                                                                             \
                out[i * n + j] = (uint8_t)in;                                                                 \
            }                                                                                                 \
        }                                                                                                     \
    }
ORC_CROP_BODY(float, f32, sinf, cosf)
ORC_CROP_BODY(double, f64, sin, cos)

/* ---- signed point-to-rotated-box distance: d3d/box/dist.cpp:11-47 pdist2dr_forward.  dist is T[m boxes, n points]
 * (positive inside), iedge u8[m, n] the edge (or vertex) that realises it.  Vertices as dgal::poly2_from_xywhr (geometry.hpp:417-429);
 * distance(Poly2, Point2, idx) geometry.hpp:482-497 over distance(Segment2, Point2) :453-474 with the line of :331-336 and the
 * projection parameter t_from_pxy :372-380 (three branches, kept as written). */
#define ORC_PDIST_BODY(T, SFX, SIN, COS, HYPOT, ABS)                                                          \
    static T orc_tpxy_##SFX(T a, T b, T c, T x, T y)                                                          \
    {                                                                                                         \
        if (b == 0) return (1 - y) / a;                                                                       \
        else if (a == 0) return (x - 1) / b;                                                                  \
        else return (b * x - a * y - a * (a + c) / b - b) / (a * a + b * b);                                  \
    }                                                                                                         \
    static T orc_segdist_##SFX(T x1, T y1, T x2, T y2, T px, T py)                                            \
    {                                                                                                         \
        const T a = y2 - y1, b = x1 - x2, c = x2 * y1 - x1 * y2;                                              \
        const T t = orc_tpxy_##SFX(a, b, c, px, py);                                                          \
        const T sign = a * px + b * py + c;                                                                   \
        if (t < orc_tpxy_##SFX(a, b, c, x2, y2)) { const T d = HYPOT(px - x2, py - y2); return sign > 0 ? d : -d; } \
        else if (t > orc_tpxy_##SFX(a, b, c, x1, y1)) { const T d = HYPOT(px - x1, py - y1); return sign > 0 ? d : -d; } \
        else return sign / HYPOT(a, b);                                                                       \
    }                                                                                                         \
    void orc_pdist2dr_##SFX(const T *pts, int64_t n, const T *boxes, int64_t m, T *dist, uint8_t *iedge)      \
    {                                                                                                         \
        for (int64_t i = 0; i < m; i++) {                                                                     \
            const T x = boxes[5 * i], y = boxes[5 * i + 1], w = boxes[5 * i + 2], h = boxes[5 * i + 3], r = boxes[5 * i + 4]; \
            const T dxsin = w * SIN(r) / 2, dxcos = w * COS(r) / 2, dysin = h * SIN(r) / 2, dycos = h * COS(r) / 2; \
            const T vx[4] = {x - dxcos + dysin, x + dxcos + dysin, x + dxcos - dysin, x - dxcos - dysin};       \
            const T vy[4] = {y - dxsin - dycos, y + dxsin - dycos, y + dxsin + dycos, y - dxsin + dycos};       \
            for (int64_t j = 0; j < n; j++) {                                                                 \
                const T px = pts[2 * j], py = pts[2 * j + 1];                                                 \
                T dmin = -orc_segdist_##SFX(vx[3], vy[3], vx[0], vy[0], px, py);                              \
                uint8_t idx = 3;                                                                              \
                for (int k = 1; k < 4; k++) {                                                                 \
                    const T dl = -orc_segdist_##SFX(vx[k - 1], vy[k - 1], vx[k], vy[k], px, py);              \
                    if (ABS(dl) < ABS(dmin)) { dmin = dl; idx = (uint8_t)(k - 1); }                           \
                }                                                                                             \
                dist[i * n + j] = dmin;                                                                       \
                if (iedge) iedge[i * n + j] = idx;                                                            \
            }                                                                                                 \
        }                                                                                                     \
    }
ORC_PDIST_BODY(float, f32, sinf, cosf, hypotf, fabsf)
ORC_PDIST_BODY(double, f64, sin, cos, hypot, fabs)

/* ---- NMS: d3d/box/nms.cpp:9-96 nms2d_templated.
 * `order` (i64[n], descending score; computed by the caller exactly like nms.cpp:103) and `scores`
 * (copied by the caller like nms.cpp:104-105) are mutated by the soft variants.
 * cuda_score_rule != 0 applies the reference CUDA rule instead (nms_cuda.cu:223: every box with
 * score <= thr is suppressed, including rank 0) -- SURVEY.md 8(c) T4.
 * Returns the number of IoU evaluations. */
#define ORC_NMS_BODY(T, SFX, POWF, EXPF)                                                              \
    int64_t orc_nms2d_##SFX(const T *boxes, T *scores, int64_t *order, int64_t n, int iou_type,       \
                            int sup_type, float iou_thr, float score_thr, float sup_param, int alg,   \
                            int cuda_score_rule, uint8_t *suppressed)                                 \
    {                                                                                                 \
        int64_t evals = 0;                                                                            \
        memset(suppressed, 0, (size_t)n);                                                             \
        if (cuda_score_rule) {                                                                        \
            for (int64_t i = 0; i < n; i++) if (!(scores[i] > score_thr)) suppressed[i] = 1;          \
        } else {                                                                                      \
            for (int64_t _i = n - 1; _i > 0; _i--) { /* nms.cpp:22-29: rank 0 is never swept */       \
                int64_t i = order[_i];                                                                \
                if (scores[i] > score_thr) break;                                                     \
                suppressed[i] = 1;                                                                    \
            }                                                                                         \
        }                                                                                             \
        for (int64_t _i = 0; _i < n; _i++) {                                                          \
            int64_t i = order[_i];                                                                    \
            if (suppressed[i]) { if (sup_type == ORC_SUP_HARD) continue; else break; }                \
            for (int64_t _j = _i + 1; _j < n; _j++) {                                                 \
                int64_t j = order[_j];                                                                \
                if (sup_type == ORC_SUP_HARD && suppressed[j]) continue;                              \
                T iou = (iou_type == ORC_IOU_BOX) ? aabb_iou_##SFX(boxes + 5 * i, boxes + 5 * j)      \
                                                  : quad_iou_##SFX(boxes + 5 * i, boxes + 5 * j, alg, NULL); \
                evals++;                                                                              \
                if (iou > iou_thr) { /* T vs float: promoted to T, nms.cpp:53 */                      \
                    if (sup_type == ORC_SUP_HARD) suppressed[j] = 1;                                  \
                    else if (sup_type == ORC_SUP_LINEAR) {                                            \
                        scores[j] *= 1 - POWF(iou, sup_param);                                        \
                        suppressed[j] = scores[j] < score_thr;                                        \
                    } else {                                                                          \
                        scores[j] *= EXPF(-iou * iou / sup_param);                                    \
                        suppressed[j] = scores[j] < score_thr;                                        \
                    }                                                                                 \
                }                                                                                     \
            }                                                                                         \
            if (sup_type != ORC_SUP_HARD) { /* nms.cpp:74-94 insertion re-sort */                     \
                int64_t S = n - 1;                                                                    \
                while (S > _i && !suppressed[order[S]]) S--;                                          \
                for (int64_t _j = S - 1; _j > _i; _j--) {                                             \
                    int64_t j = order[_j], _k = _j + 1;                                               \
                    while (_k < S && (suppressed[j] || scores[order[_k]] > scores[j])) {              \
                        order[_k - 1] = order[_k]; _k++;                                              \
                    }                                                                                 \
                    order[_k - 1] = j;                                                                \
                }                                                                                     \
            }                                                                                         \
        }                                                                                             \
        return evals;                                                                                 \
    }
/* nms.cpp:62,67: pow(scalar_t, float) / exp(scalar_t) resolve to the double overloads for double and,
 * through <cmath>'s promotion rules, to float pow/exp for float */
ORC_NMS_BODY(float, f, powf, expf)
ORC_NMS_BODY(double, d, pow, exp)

/* ------------------------------------------------------------------ voxelization */
/* open-addressing map (x,y,z) -> id; replaces std::unordered_map (d3d/voxel/voxelize.cpp:16-42).
 * Only find/insert-in-arrival-order is needed, so iteration order of the container never matters
 * except in voxelize_filter's coords copy, which is order independent. */
typedef struct { int32_t x, y, z, id; } orc_cell;
typedef struct { orc_cell *c; uint64_t mask; int64_t used; } orc_map;
static void orc_map_init(orc_map *m, int64_t cap)
{
    uint64_t s = 64; while ((int64_t)s < cap * 2) s <<= 1;
    m->c = (orc_cell *)malloc(s * sizeof(orc_cell));
    for (uint64_t i = 0; i < s; i++) m->c[i].id = -1;
    m->mask = s - 1; m->used = 0;
}
static inline uint64_t orc_hash3(int32_t x, int32_t y, int32_t z)
{
    uint64_t h = (uint64_t)(uint32_t)x * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)(uint32_t)y * 0xC2B2AE3D27D4EB4Full + (h >> 29);
    h ^= (uint64_t)(uint32_t)z * 0x165667B19E3779F9ull + (h << 7);
    return h ^ (h >> 32);
}
/* returns the cell for (x,y,z); *isnew set when it was just created (id left at -1 for the caller) */
static orc_cell *orc_map_get(orc_map *m, int32_t x, int32_t y, int32_t z, int *isnew)
{
    uint64_t i = orc_hash3(x, y, z) & m->mask;
    for (;;) {
        orc_cell *c = &m->c[i];
        if (c->id == -1) { /* empty slot; callers assign id >= 0 immediately */
            c->x = x; c->y = y; c->z = z; *isnew = 1; m->used++; return c;
        }
        if (c->x == x && c->y == y && c->z == z) { *isnew = 0; return c; }
        i = (i + 1) & m->mask;
    }
}

/* ---- dense: d3d/voxel/voxelize.cpp:45-180 voxelize_3d_dense_templated.
 * points f32[n,c]; shape i32[3]; bound f32[6] = xmin,xmax,ymin,ymax,zmin,zmax.
 * Outputs are caller-allocated for max_voxels entries: voxels f32[V,P,C] (caller zero-fills),
 * coords i64[V,3], pmask u8[V,P] (caller zero-fills: the reference leaves `false` slots
 * uninitialised, voxelize.cpp:58), npoints i32[V] (zero-filled), aggregates f32[V,C] or NULL.
 * reduction: 0 NONE 1 MEAN 2 MAX 3 MIN (voxelize.h:5).  Returns nvoxels. */
int64_t orc_voxelize_dense(const float *points, int64_t n, int64_t c, const int32_t *shape, const float *bound,
                           int32_t max_points, int32_t max_voxels, int reduction, float *voxels, int64_t *coords,
                           uint8_t *pmask, int32_t *npoints, float *aggregates)
{
    float vsize[3];
    for (int d = 0; d < 3; d++) vsize[d] = (bound[2 * d + 1] - bound[2 * d]) / shape[d];
    if (aggregates && reduction != 0)
        for (int64_t i = 0; i < (int64_t)max_voxels * c; i++)
            aggregates[i] = reduction == 1 ? 0.0f : (reduction == 2 ? -INFINITY : INFINITY);
    orc_map map; orc_map_init(&map, n < max_voxels ? n : max_voxels);
    int64_t nvox = 0;
    for (int64_t i = 0; i < n; i++) {
        int32_t cd[3]; int oor = 0;
        for (int d = 0; d < 3; d++) {
            int idx = (int)((points[i * c + d] - bound[2 * d]) / vsize[d]); /* C truncation, :100 */
            if (idx < 0 || idx >= shape[d]) { oor = 1; break; }
            cd[d] = idx;
        }
        if (oor) continue;
        /* find first (so that a refused new voxel does not occupy a map cell) */
        int isnew; int64_t vid;
        if (nvox >= max_voxels) {
            /* probe without inserting */
            uint64_t h = orc_hash3(cd[0], cd[1], cd[2]) & map.mask; vid = -1;
            for (;;) {
                orc_cell *q = &map.c[h];
                if (q->id == -1) break;
                if (q->x == cd[0] && q->y == cd[1] && q->z == cd[2]) { vid = q->id; break; }
                h = (h + 1) & map.mask;
            }
            if (vid < 0) continue; /* :118-119 new voxel refused */
        } else {
            orc_cell *cell = orc_map_get(&map, cd[0], cd[1], cd[2], &isnew);
            if (isnew) {
                cell->id = (int32_t)nvox;
                for (int d = 0; d < 3; d++) coords[nvox * 3 + d] = cd[d];
                nvox++;
            }
            vid = cell->id;
        }
        int32_t k = npoints[vid]++;
        if (k < max_points) {
            pmask[vid * max_points + k] = 1;
            for (int64_t d = 0; d < c; d++) voxels[(vid * max_points + k) * c + d] = points[i * c + d];
        }
        if (aggregates && reduction != 0)
            for (int64_t d = 0; d < c; d++) {
                float *a = &aggregates[vid * c + d], p = points[i * c + d];
                if (reduction == 1) *a += p;
                else if (reduction == 2) *a = *a > p ? *a : p; /* std::max(a, p) */
                else *a = p < *a ? p : *a;                     /* std::min(a, p) */
            }
    }
    if (aggregates && reduction == 1)
        for (int64_t v = 0; v < nvox; v++)
            for (int64_t d = 0; d < c; d++) aggregates[v * c + d] /= npoints[v];
    free(map.c);
    return nvox;
}

/* ---- sparse: d3d/voxel/voxelize.cpp:288-335 voxelize_sparse (bound as voxelize_3d_sparse).
 * coord = (int)floor(p/size) on an unbounded grid, ids by first appearance.
 * mapping i64[n], coords i64[<=n,3], npoints i32[<=n].  Returns nvoxels. */
int64_t orc_voxelize_sparse(const float *points, int64_t n, int64_t c, const float *vsize, int64_t *mapping,
                            int64_t *coords, int32_t *npoints)
{
    orc_map map; orc_map_init(&map, n);
    int64_t nvox = 0;
    for (int64_t i = 0; i < n; i++) {
        int32_t cd[3];
        for (int d = 0; d < 3; d++) cd[d] = (int32_t)floorf(points[i * c + d] / vsize[d]); /* :309 */
        int isnew; orc_cell *cell = orc_map_get(&map, cd[0], cd[1], cd[2], &isnew);
        if (isnew) {
            cell->id = (int32_t)nvox;
            for (int d = 0; d < 3; d++) coords[nvox * 3 + d] = cd[d];
            npoints[nvox] = 1; nvox++;
        } else npoints[cell->id] += 1;
        mapping[i] = cell->id;
    }
    free(map.c);
    return nvox;
}

static int orc_cmp_desc_stable(const void *a, const void *b)
{
    const int64_t *x = (const int64_t *)a, *y = (const int64_t *)b; /* {count, id} */
    if (x[0] != y[0]) return x[0] > y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1]);
}

/* ---- filter: d3d/voxel/voxelize.cpp:337-484 voxelize_filter.
 * in: mapping i64[n], coords i64[nv,3], npoints i32[nv], bound i64[3,2] or NULL.
 * pfilter 0 NONE 1 TRIM; vfilter 0 NONE 1 TRIM 2 DESCENDING (voxelize.h:6-7).
 * DESCENDING visits voxels by descending count; ties broken by ascending id here (the reference's
 * torch.argsort leaves ties unspecified).
 * out: out_mask i64[<=n] (surviving input indices), out_mapping i64[<=n], out_npoints i32[<=nv],
 * out_coords i64[<=nv,3].  Returns kept points in *k_out and kept voxels as the return value. */
int64_t orc_voxelize_filter(int64_t n, const int64_t *mapping, const int64_t *coords, const int32_t *npoints,
                            int64_t nv, const int64_t *bound, int32_t min_points, int32_t max_points,
                            int32_t max_voxels, int pfilter, int vfilter, int64_t *out_mask, int64_t *out_mapping,
                            int32_t *out_npoints, int64_t *out_coords, int64_t *k_out)
{
    int64_t *newid = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nv > 0 ? nv : 1));
    for (int64_t v = 0; v < nv; v++) newid[v] = -1;
    int64_t kept = 0;
#define OOB(v) (bound && (coords[(v)*3] < bound[0] || coords[(v)*3] >= bound[1] || coords[(v)*3 + 1] < bound[2] || \
                          coords[(v)*3 + 1] >= bound[3] || coords[(v)*3 + 2] < bound[4] || coords[(v)*3 + 2] >= bound[5]))
    if (vfilter == 0 || vfilter == 1) {
        for (int64_t v = 0; v < nv; v++) {
            if (vfilter == 1 && kept >= max_voxels) break;
            if (npoints[v] < min_points) continue;
            if (OOB(v)) continue;
            newid[v] = kept++;
        }
    } else {
        int64_t *ord = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(nv > 0 ? nv : 1));
        for (int64_t v = 0; v < nv; v++) { ord[2 * v] = npoints[v]; ord[2 * v + 1] = v; }
        qsort(ord, (size_t)nv, 2 * sizeof(int64_t), orc_cmp_desc_stable);
        for (int64_t q = 0; q < nv; q++) {
            int64_t v = ord[2 * q + 1];
            if (kept >= max_voxels) break;
            if (npoints[v] < min_points) break; /* :409-410 */
            if (OOB(v)) continue;
            newid[v] = kept++;
        }
        free(ord);
    }
#undef OOB
    for (int64_t v = 0; v < nv; v++)
        if (newid[v] >= 0) for (int d = 0; d < 3; d++) out_coords[newid[v] * 3 + d] = coords[v * 3 + d];
    for (int64_t v = 0; v < kept; v++) out_npoints[v] = 0;
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t nid = newid[mapping[i]];
        if (nid < 0) continue;
        if (pfilter == 1 && out_npoints[nid] >= max_points) continue; /* :452-456 */
        out_npoints[nid]++;
        out_mask[k] = i; out_mapping[k] = nid; k++;
    }
    *k_out = k;
    free(newid);
    return kept;
}

/* ------------------------------------------------------------------ aligned scatter */
/* d3d/point/scatter.cpp:34-77 _fill_lcoords + :79-134 forward + :136-172 backward.
 * coord T[n,1+dim] (col 0 = batch), image T[b,c,D1..Ddim]; atype 1 MEAN 2 LINEAR (scatter.h:37). */
#define ORC_SCATTER_BODY(T, SFX)                                                                       \
    static void orc_lcoords_##SFX(const int64_t *dims, int dim, const T *crd, int atype, int lc[8][3], T lw[8]) \
    {                                                                                                  \
        int nb = 1 << dim;                                                                             \
        for (int j = 0; j < nb; j++) lw[j] = 1;                                                        \
        for (int j = 0; j < nb; j++)                                                                   \
            for (int d = 0; d < dim; d++) {                                                            \
                int dmax = (int)dims[d] - 1; T x = crd[d + 1];                                         \
                if (x > dmax) { lc[j][d] = dmax; if (atype == 2) lw[j] *= 0.5; }                       \
                else if (x < 0) { lc[j][d] = 0; if (atype == 2) lw[j] *= 0.5; }                        \
                else if (j & (1u << d)) {                                                              \
                    int k = (int)x; if (k < x) k++; /* _ceil, scatter.cpp:28-33 */                     \
                    lc[j][d] = k; if (atype == 2) lw[j] *= 1 + x - k;                                  \
                } else {                                                                               \
                    int k = (int)x; if (k > x) k--; /* _floor, scatter.cpp:22-27 */                    \
                    lc[j][d] = k; if (atype == 2) lw[j] *= 1 - x + k;                                  \
                }                                                                                      \
            }                                                                                          \
    }                                                                                                  \
    void orc_scatter_fwd_##SFX(const T *coord, int64_t n, int dim, const T *image, int64_t nb_, int64_t nc,   \
                               const int64_t *dims, int atype, T *out)                                 \
    {                                                                                                  \
        (void)nb_; int64_t plane = 1; for (int d = 0; d < dim; d++) plane *= dims[d];                  \
        int nb = 1 << dim;                                                                             \
        for (int64_t i = 0; i < n; i++) {                                                              \
            const T *crd = coord + i * (dim + 1); int b = (int)crd[0];                                 \
            int lc[8][3]; T lw[8]; orc_lcoords_##SFX(dims, dim, crd, atype, lc, lw);                   \
            for (int64_t c = 0; c < nc; c++) {                                                         \
                const T *pl = image + ((int64_t)b * nc + c) * plane; T sum = 0;                        \
                for (int j = 0; j < nb; j++) {                                                         \
                    int64_t off = 0; for (int d = 0; d < dim; d++) off = off * dims[d] + lc[j][d];     \
                    if (atype == 1) sum += pl[off]; else sum += pl[off] * lw[j];                       \
                }                                                                                      \
                out[i * nc + c] = atype == 1 ? sum / nb : sum;                                         \
            }                                                                                          \
        }                                                                                              \
    }                                                                                                  \
    void orc_scatter_bwd_##SFX(const T *coord, int64_t n, int dim, const T *grad, int64_t nb_, int64_t nc,    \
                               const int64_t *dims, int atype, T *image_grad)                          \
    {                                                                                                  \
        (void)nb_; int64_t plane = 1; for (int d = 0; d < dim; d++) plane *= dims[d];                  \
        int nb = 1 << dim;                                                                             \
        for (int64_t i = 0; i < n; i++) {                                                              \
            const T *crd = coord + i * (dim + 1); int b = (int)crd[0];                                 \
            int lc[8][3]; T lw[8]; orc_lcoords_##SFX(dims, dim, crd, atype, lc, lw);                   \
            for (int64_t c = 0; c < nc; c++) {                                                         \
                T *pl = image_grad + ((int64_t)b * nc + c) * plane;                                    \
                for (int j = 0; j < nb; j++) {                                                         \
                    int64_t off = 0; for (int d = 0; d < dim; d++) off = off * dims[d] + lc[j][d];     \
                    if (atype == 1) pl[off] += grad[i * nc + c] / nb;                                  \
                    else pl[off] += grad[i * nc + c] * lw[j];                                          \
                }                                                                                      \
            }                                                                                          \
        }                                                                                              \
    }
ORC_SCATTER_BODY(float, f32)
ORC_SCATTER_BODY(double, f64)

int orc_version(void) { return 1; }
