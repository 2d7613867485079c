// dgal_wrap_shim.cpp -- TEST INFRASTRUCTURE: exports the reference's own scalar 3-D box routines (d3d/dgal_wrap.h:6-91, compiled
// unmodified from /root/reference where it lies) behind a C ABI, so that the oracle's restatement of the detection-evaluation
// distance (oracle.box3d_iou_distance) and of box3dr_pdist is pinned to the reference itself.  Built by oracle/build_ref.py into
// oracle/_ref/libdgal_wrap.so; nothing under d3d_b200/ may load it.  The includes in front of the header are SURVEY.md finding F6
// (geometry.hpp calls unqualified abs() and assert without including their headers).
#include <math.h>
#include <stdlib.h>
#include <cassert>
#include <cmath>
#include <stdint.h>
using std::abs;
#include "d3d/dgal_wrap.h"

extern "C" {
// the pair loops of ScoreMatcher.prepare_boxes (d3d/tracking/matcher.pyx:55-76): dist[i][j] = 1 - box3d[r]_iou(src i, dst j), float32;
// boxes are [n,7] rows (x, y, z, lx, ly, lz, rz), already clipped by the caller like matcher.pyx:50-52
void ref_iou3d_distance(const float *a, int64_t n, const float *b, int64_t m, int rotated, float *dist)
{
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) {
            const float *p = a + 7 * i, *q = b + 7 * j;
            const float v = rotated ? box3dr_iou(p[0], p[1], p[2], p[3], p[4], p[5], p[6], q[0], q[1], q[2], q[3], q[4], q[5], q[6])
                                    : box3d_iou(p[0], p[1], p[2], p[3], p[4], p[5], p[6], q[0], q[1], q[2], q[3], q[4], q[5], q[6]);
            dist[i * m + j] = 1 - v;
        }
}
// abstraction.pyx:328-343: signed distance from points [n,3] to one 3-D box
void ref_box3dr_pdist(const float *box, const float *pts, int64_t n, float *out)
{
    for (int64_t j = 0; j < n; j++)
        out[j] = box3dr_pdist(box[0], box[1], box[2], box[3], box[4], box[5], box[6], pts[3 * j], pts[3 * j + 1], pts[3 * j + 2]);
}
}
