// voxel_tiles.cu -- sparse voxelization as a software pipeline of streaming tiles (the default fast path).
//
// Replaces voxelize_sparse + voxelize_filter (reference d3d/voxel/voxelize.cpp:288-484) for the common
// configurations (no voxel cap, max_points filter NONE or TRIM with max_points <= 8); everything else goes to
// the cluster path (voxel_cluster.cu) or the sort path (voxel.cu).  Same packed outputs, bit for bit.
//
// Why not one cluster per frame (voxel_cluster.cu): that kernel keeps a whole frame in the shared memory of 8 SMs,
// which pins one 1024-thread CTA per SM, all warps of an SM in the same phase, on the 120 SMs that 8-CTA clusters
// reach -- it is bound by exposed latency, not by HBM (ncu: issue slots 52 % busy, DRAM 15 %).  Here every stage is
// an ordinary tile of 256 threads with up to 7 CTAs per SM on all 148 SMs, and the stages of different frames run
// side by side in one launch ("tick"):
//
//   split   (1024 points / CTA)  cell keys with the reference's fp32 arithmetic; a point's (key, index) goes to
//                                the queue of the BUCKET its key hashes to (128 buckets for a 120k-point frame);
//                                the tile reserves its share of every queue with one atomic per bucket.
//   bucket  (one CTA / bucket)   all points of a voxel meet in one bucket: hash table in shared memory (CAS claim,
//                                atomicMin of the point index, count), the K smallest indices of voxels with more
//                                than K points by a min-cascade (slot j keeps the j-th smallest of everything it
//                                is offered and passes the larger value on: the result does not depend on the
//                                arrival order), one reply word to every point that shares its voxel.
//   write   (1024 points / CTA)  first-of-voxel and keep bits -> ballots -> one chained scan over all tiles of all
//                                frames (decoupled look-back) gives voxel ids in order of first appearance and
//                                packed rows; voxel rows and kept point rows are streamed out.
//
// Tick k launches split(chunk k) + bucket(chunk k-1) + write(chunk k-2), a chunk being a few frames, so the
// scratch of the frames in flight (queues, reply words, keys: ~1.7 MB per frame, plus the frame itself) lives
// in L2 and HBM sees the algorithmic traffic only: 16 B/point in, 32 B/kept point + 28 B/voxel out.
//
// Determinism: every output is a function of the point set (smallest index, count, K smallest indices, prefix
// sums in point order), never of the race order of the queues and tables.
//
// Anything that does not fit (a bucket queue or table overflow: e.g. tens of thousands of points in one voxel)
// raises a device flag; the cluster kernel is launched behind the pipeline on that flag and redoes the batch.
#include "voxel.cuh"
#include <stdlib.h>

namespace d3d {

constexpr int VT_THREADS = 256;
constexpr int VT_WARPS = VT_THREADS / 32;
constexpr int VT_PPT = 4;                         // points per thread
constexpr int VT_TILE = VT_THREADS * VT_PPT;      // 1024 points per tile
constexpr int VT_ROWS = VT_TILE / 32;             // 32-point rows per tile
constexpr int VT_STAGES = 3;
constexpr int VT_MAXK = 8;                        // deepest min-cascade (max_points of the TRIM filter)
constexpr int VT_POOL = 128;                      // crowded-voxel records per bucket
constexpr uint32_t VT_NONE = 0xffffffffu;
constexpr uint32_t VT_HEAD = 1u << 31, VT_KEEP = 1u << 30, VT_VAL = (1u << 30) - 1;   // reply word
constexpr unsigned long long VT_PVAL = (1ull << 62) - 1;                               // status word: flag << 62 | kept rows << 31 | voxel rows
constexpr uint32_t VT_F31 = 0x7fffffffu;

struct VtGeom {
    uint32_t lmax, lpad, tpf;      // longest frame, padded to whole tiles, tiles per frame
    uint32_t lgP, P, lgS, S, qcap; // buckets per frame, table slots per bucket, queue entries per bucket
    uint32_t CF, nchunks;          // frames per chunk, chunks
};

struct VtArgs {
    const float *pts; int nfeat; const int64_t *offs; int64_t nframes;
    float *out_points; int64_t *out_mask; int64_t *out_mapping; int32_t *out_npoints; int64_t *out_coords; int64_t *counts;
    uint2 *queue; uint32_t *reply, *keyarr; uint4 *rowinfo; uint32_t *qcount;
    unsigned long long *status; uint32_t *tickets, *bail;
    VtGeom g;
    uint32_t K, cthr, dflt; int min_points; int trim, dropall;
};

__device__ __forceinline__ uint32_t vt_h(unsigned long long x) { return (uint32_t)x & VT_F31; }
__device__ __forceinline__ uint32_t vt_k(unsigned long long x) { return (uint32_t)(x >> 31) & VT_F31; }
__device__ __forceinline__ unsigned long long vt_pack(uint32_t h, uint32_t k) { return ((unsigned long long)k << 31) | h; }

__device__ __forceinline__ unsigned long long vt_ld_acquire(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void vt_st_release(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Decoupled look-back over the tiles of all frames (one warp).  Tiles take their numbers from a ticket counter, so a
// tile only waits for tiles that are running or done.  flag 1: the tile's own sums; flag 2: inclusive prefix; flag 3:
// inclusive prefix and the tile's row table is published.
__device__ __forceinline__ unsigned long long vt_lookback(unsigned long long *state, int64_t gt, unsigned long long mine)
{
    const unsigned lane = threadIdx.x & 31u;
    if (gt > 0 && lane == 0) vt_st_release(state + gt, (1ull << 62) | mine);
    unsigned long long excl = 0;
    for (int64_t j = gt - 1; j >= 0; j -= 32) {
        const int64_t idx = j - (int64_t)lane;
        unsigned long long v = 2ull << 62;   // before the first tile: an inclusive prefix of zero
        unsigned first2, need;
        for (;;) {
            if (idx >= 0) v = vt_ld_acquire(state + idx);
            const unsigned flag = (unsigned)(v >> 62);
            const unsigned b2 = __ballot_sync(0xffffffffu, flag >= 2), b0 = __ballot_sync(0xffffffffu, flag == 0);
            first2 = b2 ? (unsigned)__ffs((int)b2) - 1u : 32u;
            need = first2 >= 31u ? 0xffffffffu : ((2u << first2) - 1u);
            if (!(b0 & need)) break;
            __nanosleep(40);
        }
        unsigned long long x = ((need >> lane) & 1u) ? (v & VT_PVAL) : 0ull;
#pragma unroll
        for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        excl += x;
        if (first2 < 32u) break;
    }
    if (lane == 0) vt_st_release(state + gt, (2ull << 62) | (excl + mine));
    return excl;
}

__device__ __forceinline__ uint32_t vt_bucket(uint32_t key, uint32_t lgP) { return lgP ? (key * 0x9E3779B1u) >> (32u - lgP) : 0u; }
__device__ __forceinline__ uint32_t vt_home(uint32_t key, uint32_t lgS) { return (key * 0x85EBCA6Bu) >> (32u - lgS); }

// ------------------------------------------------------------------------------------------------ split
__device__ __forceinline__ void vt_split(const VtArgs &a, const VcDev &dv, uint32_t lt, uint32_t chunk, unsigned char *dyn)
{
    const VtGeom &g = a.g;
    const unsigned tid = threadIdx.x;
    const uint32_t fl = lt / g.tpf, t = lt - fl * g.tpf;
    const int64_t f = (int64_t)chunk * g.CF + fl;
    const uint32_t slot = (chunk % VT_STAGES) * g.CF + fl;
    const int64_t b = a.offs[f];
    const uint32_t L = (uint32_t)min((long long)(a.offs[f + 1] - b), (long long)g.lmax);
    const uint32_t t0 = t * VT_TILE;
    if (t0 >= L) return;

    uint32_t *hist = reinterpret_cast<uint32_t *>(dyn), *base = hist + g.P;
    uint32_t *reply = a.reply + (size_t)slot * g.lpad, *keyarr = a.keyarr + (size_t)slot * g.lpad;
    uint32_t *qcount = a.qcount + (size_t)slot * g.P;
    uint2 *queue = a.queue + (size_t)slot * g.P * g.qcap;

    for (uint32_t p = tid; p < g.P; p += VT_THREADS) hist[p] = 0;
    __syncthreads();

    float4 p4[VT_PPT];
    bool in[VT_PPT];
#pragma unroll
    for (int u = 0; u < VT_PPT; u++) {
        const uint32_t i = t0 + u * VT_THREADS + tid;
        in[u] = i < L;
        p4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in[u]) {
            if (a.nfeat == 4) p4[u] = __ldg(reinterpret_cast<const float4 *>(a.pts) + b + i);
            else { const float *q = a.pts + (b + i) * a.nfeat; p4[u] = make_float4(q[0], q[1], q[2], 0.f); }
        }
    }
    uint32_t keys[VT_PPT], rank[VT_PPT];
#pragma unroll
    for (int u = 0; u < VT_PPT; u++) {
        const uint32_t i = t0 + u * VT_THREADS + tid;
        uint32_t key;
        const bool ok = vc_cell<false>(dv, p4[u], &key) && in[u];
        keys[u] = ok ? key : VC_NOKEY;
        rank[u] = 0;
        if (in[u]) { keyarr[i] = keys[u]; reply[i] = ok ? a.dflt : VT_NONE; }
        if (ok) rank[u] = atomicAdd(&hist[vt_bucket(key, g.lgP)], 1u);
    }
    __syncthreads();
    for (uint32_t p = tid; p < g.P; p += VT_THREADS) {
        const uint32_t c = hist[p];
        base[p] = c ? atomicAdd(&qcount[p], c) : 0u;
    }
    __syncthreads();
    bool over = false;
#pragma unroll
    for (int u = 0; u < VT_PPT; u++) {
        if (keys[u] != VC_NOKEY) {
            const uint32_t bk = vt_bucket(keys[u], g.lgP);
            const uint32_t pos = base[bk] + rank[u];
            if (pos < g.qcap) queue[(size_t)bk * g.qcap + pos] = make_uint2(keys[u], t0 + u * VT_THREADS + tid);
            else over = true;
        }
    }
    if (over) *a.bail = 1u;
}

// ------------------------------------------------------------------------------------------------ bucket
__device__ __forceinline__ void vt_bucket_role(const VtArgs &a, uint32_t lb, uint32_t chunk, unsigned char *dyn)
{
    const VtGeom &g = a.g;
    const unsigned tid = threadIdx.x;
    const uint32_t fl = lb >> g.lgP, bk = lb & (g.P - 1);
    const uint32_t slot = (chunk % VT_STAGES) * g.CF + fl;
    const uint32_t S = g.S, smask = S - 1, K = a.K, cthr = a.cthr;

    uint32_t *tkey = reinterpret_cast<uint32_t *>(dyn), *tmin = tkey + S, *tcnt = tmin + S, *taux = tcnt + S;
    uint32_t *pool = taux + S;                       // VT_POOL records of VT_MAXK indices
    uint32_t *misc = pool + VT_POOL * VT_MAXK;       // [0] n, [1] records in use, [2] failure, [3] the batch already failed elsewhere
    uint32_t *qcount = a.qcount + (size_t)slot * g.P;
    const uint2 *q = a.queue + ((size_t)slot * g.P + bk) * g.qcap;
    uint32_t *reply = a.reply + (size_t)slot * g.lpad;

    if (tid == 0) { misc[0] = min(qcount[bk], g.qcap); misc[1] = 0; misc[2] = 0; misc[3] = *reinterpret_cast<volatile uint32_t *>(a.bail); qcount[bk] = 0; }   // the counter is ready for the slot's next frame
    for (uint32_t s = tid; s < S; s += VT_THREADS) { tkey[s] = VT_NONE; tmin[s] = VT_NONE; tcnt[s] = 0; }
    __syncthreads();
    const uint32_t n = misc[0];
    if (n == 0 || misc[3]) return;

    // a: claim a slot per key (linear probing), smallest index and point count per voxel
    for (uint32_t e = tid; e < n; e += VT_THREADS) {
        const uint2 en = q[e];
        uint32_t s = vt_home(en.x, g.lgS), it = 0;
        for (;;) {
            const uint32_t old = atomicCAS(&tkey[s], VT_NONE, en.x);
            if (old == VT_NONE || old == en.x) break;
            s = (s + 1) & smask;
            if (++it >= S) break;
        }
        if (it >= S) misc[2] = 1u;   // more voxels than slots
        else { atomicMin(&tmin[s], en.y); atomicAdd(&tcnt[s], 1u); }
    }
    __syncthreads();
    if (misc[2]) { if (tid == 0) *a.bail = 1u; return; }

    auto find = [&](uint32_t key) -> uint32_t {
        uint32_t s = vt_home(key, g.lgS);
        while (tkey[s] != key) s = (s + 1) & smask;
        return s;
    };

    // b: voxels with more than K points -- their K smallest indices
    if (cthr != VT_NONE) {
        bool anyc = false;
        for (uint32_t e = tid; e < n; e += VT_THREADS) {
            const uint2 en = q[e];
            const uint32_t s = find(en.x);
            if (tcnt[s] > cthr) {
                anyc = true;
                if (tmin[s] == en.y) {   // the voxel's first point opens the record
                    const uint32_t r = atomicAdd(&misc[1], 1u);
                    if (r >= (uint32_t)VT_POOL) misc[2] = 1u;
                    else {
                        taux[s] = r;
                        for (uint32_t j = 0; j < K; j++) pool[r * VT_MAXK + j] = VT_NONE;
                    }
                }
            }
        }
        if (__syncthreads_or((int)anyc)) {
            if (misc[2]) { if (tid == 0) *a.bail = 1u; return; }
            for (uint32_t e = tid; e < n; e += VT_THREADS) {
                const uint2 en = q[e];
                const uint32_t s = find(en.x);
                if (tcnt[s] > cthr) {
                    uint32_t *rec = pool + taux[s] * VT_MAXK;
                    uint32_t x = en.y;
                    for (uint32_t j = 0; j < K; j++) {
                        const uint32_t old = atomicMin(&rec[j], x);
                        x = max(old, x);              // the larger value moves on to the next level
                        if (x == VT_NONE) break;
                    }
                }
            }
            __syncthreads();
        }
    }

    // c: one word to every point that shares its voxel (a point that hears nothing is the only point of its voxel)
    for (uint32_t e = tid; e < n; e += VT_THREADS) {
        const uint2 en = q[e];
        const uint32_t s = find(en.x);
        const uint32_t total = tcnt[s];
        if (total == 1) continue;
        const uint32_t mn = tmin[s];
        uint32_t r;
        if ((long long)total < (long long)a.min_points) r = VT_NONE;
        else if (en.y == mn) r = VT_HEAD | VT_KEEP | min(total, VT_VAL);
        else {
            const bool kept = !(total > cthr) || en.y <= pool[taux[s] * VT_MAXK + K - 1];
            r = (kept ? VT_KEEP : 0u) | mn;
        }
        reply[en.y] = r;
    }
}

// ------------------------------------------------------------------------------------------------ write
__device__ __forceinline__ void vt_write(const VtArgs &a, const VcDev &dv, uint32_t chunk, unsigned char *dyn)
{
    const VtGeom &g = a.g;
    const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const unsigned ltmask = lanemask_lt();

    unsigned long long *wsum = reinterpret_cast<unsigned long long *>(dyn);   // [8] warp sums, then exclusive warp bases
    unsigned long long *s_excl = wsum + VT_WARPS;                             // [1]
    uint32_t *s_misc = reinterpret_cast<uint32_t *>(s_excl + 1);              // [0] ticket, [1] first voxel row of the frame
    uint4 *ri_s = reinterpret_cast<uint4 *>(dyn + 128);                       // [VT_ROWS] this tile's row table
    long long *stg = reinterpret_cast<long long *>(dyn + 128 + VT_ROWS * 16) + w * 96;   // per-warp staging of 32 voxel rows

    if (tid == 0) s_misc[0] = atomicAdd(&a.tickets[chunk], 1u);
    __syncthreads();
    const uint32_t lt = s_misc[0];
    const uint32_t fl = lt / g.tpf, t = lt - fl * g.tpf;
    const int64_t f = (int64_t)chunk * g.CF + fl;
    const uint32_t slot = (chunk % VT_STAGES) * g.CF + fl;
    const int64_t gt = f * (int64_t)g.tpf + t;
    const int64_t b = a.offs[f];
    const uint32_t L = (uint32_t)min((long long)(a.offs[f + 1] - b), (long long)g.lmax);
    const uint32_t t0 = t * VT_TILE;
    const uint32_t *reply = a.reply + (size_t)slot * g.lpad, *keyarr = a.keyarr + (size_t)slot * g.lpad;
    uint4 *ri_f = a.rowinfo + (size_t)slot * (g.lpad / 32);
    unsigned long long *fstatus = a.status + f * (int64_t)g.tpf;

    uint32_t rv[VT_PPT], hb[VT_PPT], kb[VT_PPT];
    uint32_t hs = 0, ks = 0;
#pragma unroll
    for (int u = 0; u < VT_PPT; u++) {
        const uint32_t i = t0 + (w * VT_PPT + u) * 32 + lane;
        uint32_t r = i < L ? reply[i] : VT_NONE;
        if (r != VT_NONE && a.dropall) r &= ~VT_KEEP;
        if (r != VT_NONE && !(r & (VT_HEAD | VT_KEEP))) r = VT_NONE;   // a dropped point that is not the first of its voxel
        rv[u] = r;
        hb[u] = __ballot_sync(0xffffffffu, r != VT_NONE && (r & VT_HEAD));
        kb[u] = __ballot_sync(0xffffffffu, r != VT_NONE && (r & VT_KEEP));
        hs += __popc(hb[u]); ks += __popc(kb[u]);
    }
    if (lane == 0) wsum[w] = vt_pack(hs, ks);
    __syncthreads();
    if (w == 0) {
        const unsigned long long v = lane < (unsigned)VT_WARPS ? wsum[lane] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < VT_WARPS; d <<= 1) { const unsigned long long x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += x; }
        const unsigned long long total = __shfl_sync(0xffffffffu, inc, VT_WARPS - 1);
        const unsigned long long ex = vt_lookback(a.status, gt, total);
        if (lane < (unsigned)VT_WARPS) wsum[lane] = inc - v;
        if (lane == 0) {
            *s_excl = ex;
            uint32_t vb;
            if (t == 0) {
                vb = vt_h(ex);
                a.counts[2 * f] = (long long)vt_k(ex); a.counts[2 * f + 1] = (long long)vt_h(ex);
            } else {   // the frame's first voxel row: the first row of the frame's first tile (an earlier ticket)
                while ((vt_ld_acquire(fstatus) >> 62) != 3ull) __nanosleep(40);
                vb = __ldcg(reinterpret_cast<const uint32_t *>(ri_f) + 1);
            }
            s_misc[1] = vb;
            if (gt == a.nframes * (int64_t)g.tpf - 1) {
                a.counts[2 * a.nframes] = (long long)vt_k(ex + total); a.counts[2 * a.nframes + 1] = (long long)vt_h(ex + total);
            }
        }
    }
    __syncthreads();
    uint32_t Gh[VT_PPT], Gk[VT_PPT];
    {
        unsigned long long run = *s_excl + wsum[w];
#pragma unroll
        for (int u = 0; u < VT_PPT; u++) {
            Gh[u] = vt_h(run); Gk[u] = vt_k(run);
            if (lane == 0) {
                const uint4 ri = make_uint4(hb[u], Gh[u], kb[u], Gk[u]);
                ri_s[w * VT_PPT + u] = ri;
                __stcg(ri_f + t * VT_ROWS + w * VT_PPT + u, ri);
            }
            run += vt_pack(__popc(hb[u]), __popc(kb[u]));
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) vt_st_release(a.status + gt, (3ull << 62) | (vt_ld_acquire(a.status + gt) & VT_PVAL));   // row table published
    const uint32_t vb = s_misc[1];

    // point and key loads of the four rows in flight together
    float4 p4[VT_PPT];
    uint32_t ckey[VT_PPT];
#pragma unroll
    for (int u = 0; u < VT_PPT; u++) {
        const uint32_t i = t0 + (w * VT_PPT + u) * 32 + lane;
        p4[u] = make_float4(0.f, 0.f, 0.f, 0.f); ckey[u] = 0;
        if (rv[u] != VT_NONE) {
            if ((rv[u] & VT_KEEP) && a.nfeat == 4) p4[u] = __ldg(reinterpret_cast<const float4 *>(a.pts) + b + i);
            if (rv[u] & VT_HEAD) ckey[u] = keyarr[i];
        }
    }
    const uint32_t mask_y = (1u << (dv.sh_x - dv.sh_y)) - 1u, mask_z = (1u << dv.sh_y) - 1u;
#pragma unroll
    for (int u = 0; u < VT_PPT; u++) {
        const uint32_t row = w * VT_PPT + u;
        const uint32_t i = t0 + row * 32 + lane;
        const uint32_t r = rv[u];
        const bool valid = r != VT_NONE;
        const bool head = valid && (r & VT_HEAD), keep = valid && (r & VT_KEEP);
        uint32_t vid = 0;
        if (head) vid = Gh[u] + __popc(hb[u] & ltmask) - vb;
        else if (keep) {   // the id lives where the voxel's first point lives
            const uint32_t m = r & VT_VAL, mt = m / VT_TILE;
            uint4 rj;
            if (mt == t) rj = ri_s[(m >> 5) & (VT_ROWS - 1)];
            else {
                while ((vt_ld_acquire(fstatus + mt) >> 62) != 3ull) __nanosleep(40);
                rj = __ldcg(ri_f + (m >> 5));
            }
            vid = rj.y + __popc(rj.x & ((1u << (m & 31u)) - 1u)) - vb;
        }
        // voxel rows of this 32-point row are consecutive: the coordinates go through a per-warp staging row so that the
        // global stores are dense lines instead of 24-byte-strided ones
        if (hb[u]) {
            const uint32_t rank = __popc(hb[u] & ltmask), nh3 = 3u * __popc(hb[u]);
            if (head) {
                const uint32_t c = r & VT_VAL, key = ckey[u];
                stg[rank * 3 + 0] = (long long)(key >> dv.sh_x) + dv.cadd[0];
                stg[rank * 3 + 1] = (long long)((key >> dv.sh_y) & mask_y) + dv.cadd[1];
                stg[rank * 3 + 2] = (long long)(key & mask_z) + dv.cadd[2];
                __stcs(a.out_npoints + Gh[u] + rank, (a.trim && c > a.K) ? (int32_t)a.K : (int32_t)c);
            }
            __syncwarp();
            long long *co = reinterpret_cast<long long *>(a.out_coords) + (size_t)Gh[u] * 3;
#pragma unroll
            for (int q = 0; q < 3; q++)
                if (q * 32 + lane < nh3) __stcs(co + q * 32 + lane, stg[q * 32 + lane]);
            __syncwarp();
        }
        if (keep) {
            const size_t o = (size_t)Gk[u] + __popc(kb[u] & ltmask);
            if (a.nfeat == 4) __stcs(reinterpret_cast<float4 *>(a.out_points) + o, p4[u]);
            else for (int q = 0; q < a.nfeat; q++) a.out_points[o * a.nfeat + q] = a.pts[(b + i) * a.nfeat + q];
            __stcs(reinterpret_cast<long long *>(a.out_mask) + o, (long long)i);
            __stcs(reinterpret_cast<long long *>(a.out_mapping) + o, (long long)vid);
        }
    }
}

// one tick of the pipeline: blocks [0, n3) write chunk c3, blocks [n3, n3 + n2) resolve the buckets of chunk c2, the
// rest split chunk c1.  The write tiles come first: they carry the chained scan.
__global__ void __launch_bounds__(VT_THREADS, 5) vt_tick_kernel(const VtArgs a, const VcDev dv, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t n3, uint32_t n2)
{
    extern __shared__ __align__(16) unsigned char vt_dyn[];
    const uint32_t bid = blockIdx.x;
    if (bid < n3) vt_write(a, dv, c3, vt_dyn);
    else if (bid < n3 + n2) vt_bucket_role(a, bid - n3, c2, vt_dyn);
    else vt_split(a, dv, bid - n3 - n2, c1, vt_dyn);
}

// ------------------------------------------------------------------------------------------------ host side
static uint32_t vt_lg2_ceil(uint64_t x) { uint32_t l = 0; while ((1ull << l) < x) l++; return l; }

static int vt_env_cf()
{
    static int cf = -1;   // read once: the environment is a tuning aid, not part of the call
    if (cf < 0) { const char *e = getenv("D3D_B200_VOX_CF"); int v = e ? atoi(e) : 0; cf = v > 0 ? v : 0; }
    return cf;
}

static bool vt_geom(int64_t max_frame_points, int64_t nframes, VtGeom *g)
{
    if (max_frame_points < 1) max_frame_points = 1;
    if (max_frame_points > (1ll << 21)) return false;            // 2048 buckets of 1024 points
    g->lmax = (uint32_t)max_frame_points;
    g->tpf = (g->lmax + VT_TILE - 1) / VT_TILE;
    g->lpad = g->tpf * VT_TILE;
    g->lgP = vt_lg2_ceil(((uint64_t)g->lmax + 1023) / 1024);
    g->P = 1u << g->lgP;
    const uint32_t per = (g->lmax + g->P - 1) / g->P;            // points per bucket if every point is kept and the hash is even
    g->lgS = vt_lg2_ceil(2ull * per);
    if (g->lgS < 6) g->lgS = 6;
    if (g->lgS > 11) return false;                                // cannot happen with per <= 1024
    g->S = 1u << g->lgS;
    g->qcap = (4 * per + 512 + 3) & ~3u;
    int cf = vt_env_cf();
    if (cf <= 0) cf = 8;
    if ((int64_t)cf > nframes) cf = (int)(nframes > 0 ? nframes : 1);
    g->CF = (uint32_t)cf;
    g->nchunks = (uint32_t)((nframes + cf - 1) / cf);
    if ((int64_t)nframes * g->tpf >= (1ll << 31)) return false;
    return true;
}

struct VtLayout { size_t queue, reply, keyarr, rowinfo, zero, qcount, status, tickets, bail, zero_end, total; };

static VtLayout vt_layout(const VtGeom &g, int64_t nframes)
{
    VtLayout l;
    const size_t slots = (size_t)VT_STAGES * g.CF;
    size_t o = 0;
    l.queue = o;   o += align_up(slots * g.P * g.qcap * sizeof(uint2));
    l.reply = o;   o += align_up(slots * g.lpad * 4);
    l.keyarr = o;  o += align_up(slots * g.lpad * 4);
    l.rowinfo = o; o += align_up(slots * (g.lpad / 32) * sizeof(uint4));
    l.zero = o;
    l.qcount = o;  o += align_up(slots * g.P * 4);
    l.status = o;  o += align_up((size_t)(nframes > 0 ? nframes : 1) * g.tpf * 8);
    l.tickets = o; o += align_up((size_t)g.nchunks * 4);
    l.bail = o;    o += 256;
    l.zero_end = o;
    l.total = o;
    return l;
}

bool vox_tiles_supported(const VoxCfg &cfg, int64_t total, int64_t nframes, int64_t max_frame_points)
{
    (void)total;
    VcDev d;
    VtGeom g;
    if (cfg.dense || !vc_make_dev(cfg, &d)) return false;
    if (cfg.vfilter != D3D_VF_NONE) return false;                                  // a voxel cap makes the keep decision depend on the ids
    if (cfg.pfilter == D3D_PF_TRIM && cfg.max_points > VT_MAXK) return false;
    if (cfg.pfilter != D3D_PF_TRIM && cfg.pfilter != D3D_PF_NONE) return false;
    if (!vox_cluster_supported(cfg, total, nframes, max_frame_points)) return false;   // the fallback behind the device flag
    return vt_geom(max_frame_points, nframes, &g);
}

size_t vox_tiles_ws_bytes(int64_t total, int64_t nframes, int64_t max_frame_points)
{
    if (max_frame_points <= 0 || max_frame_points > total) max_frame_points = total;
    VtGeom g;
    if (!vt_geom(max_frame_points, nframes, &g)) return 0;
    return vt_layout(g, nframes).total + 256 + vox_cluster_ws_bytes(total, nframes, max_frame_points);
}

int vox_tiles_sparse(const float *points, int64_t total, int nfeat, const int64_t *offs, int64_t nframes, int64_t max_frame_points, const VoxCfg &cfg,
                     float *out_points, int64_t *out_mask, int64_t *out_mapping, int32_t *out_npoints, int64_t *out_coords, int64_t *counts,
                     void *ws, size_t ws_bytes, cudaStream_t st)
{
    if (max_frame_points <= 0 || max_frame_points > total) max_frame_points = total;
    VtGeom g;
    VcDev dv;
    if (!vt_geom(max_frame_points, nframes, &g) || !vc_make_dev(cfg, &dv)) return D3D_ERR_UNSUPPORTED;
    const VtLayout lay = vt_layout(g, nframes);
    const size_t mine = align_up(lay.total);
    if (ws_bytes < mine + vox_cluster_ws_bytes(total > 0 ? total : 1, nframes, max_frame_points)) return D3D_ERR_WORKSPACE;
    char *w = (char *)ws;
    VtArgs a = {};
    a.pts = points; a.nfeat = nfeat; a.offs = offs; a.nframes = nframes;
    a.out_points = out_points; a.out_mask = out_mask; a.out_mapping = out_mapping; a.out_npoints = out_npoints; a.out_coords = out_coords; a.counts = counts;
    a.queue = (uint2 *)(w + lay.queue); a.reply = (uint32_t *)(w + lay.reply); a.keyarr = (uint32_t *)(w + lay.keyarr);
    a.rowinfo = (uint4 *)(w + lay.rowinfo); a.qcount = (uint32_t *)(w + lay.qcount);
    a.status = (unsigned long long *)(w + lay.status); a.tickets = (uint32_t *)(w + lay.tickets); a.bail = (uint32_t *)(w + lay.bail);
    a.g = g;
    const bool trim = cfg.pfilter == D3D_PF_TRIM;
    a.K = cfg.max_points > 0 ? (uint32_t)cfg.max_points : 0u;
    a.trim = trim ? 1 : 0;
    a.dropall = (trim && a.K == 0) ? 1 : 0;
    a.cthr = (trim && a.K > 0) ? a.K : VT_NONE;
    a.min_points = cfg.min_points;
    a.dflt = cfg.min_points <= 1 ? (VT_HEAD | VT_KEEP | 1u) : VT_NONE;

    static int smem_set[64];
    const uint32_t dyn_bucket = 16u * g.S + VT_POOL * VT_MAXK * 4u + 64u;
    const uint32_t dyn_split = 8u * g.P;
    const uint32_t dyn_write = 128u + VT_ROWS * 16u + VT_WARPS * 96u * 8u;
    uint32_t dyn = dyn_bucket > dyn_split ? dyn_bucket : dyn_split;
    if (dyn_write > dyn) dyn = dyn_write;
    int dev = 0;
    D3D_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !smem_set[dev]) {
        D3D_CUDA_TRY(cudaFuncSetAttribute(vt_tick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        smem_set[dev] = 1;
    }
    D3D_CUDA_TRY(cudaMemsetAsync(w + lay.zero, 0, lay.zero_end - lay.zero, st));
    auto nf = [&](uint32_t c) -> uint32_t { const int64_t r = nframes - (int64_t)c * g.CF; return (uint32_t)(r < (int64_t)g.CF ? r : g.CF); };
    for (uint32_t k = 0; k < g.nchunks + 2; k++) {
        const uint32_t c1 = k, c2 = k - 1, c3 = k - 2;
        const uint32_t n1 = k < g.nchunks ? nf(c1) * g.tpf : 0u;
        const uint32_t n2 = (k >= 1 && c2 < g.nchunks) ? nf(c2) * g.P : 0u;
        const uint32_t n3 = (k >= 2 && c3 < g.nchunks) ? nf(c3) * g.tpf : 0u;
        if (n1 + n2 + n3 == 0) continue;
        vt_tick_kernel<<<n1 + n2 + n3, VT_THREADS, dyn, st>>>(a, dv, c1, c2, c3, n3, n2);
        D3D_LAUNCHED();
    }
    // the batch is redone by the cluster kernel when a queue, a table or a record pool overflowed (device flag)
    return vox_cluster_sparse(points, total, nfeat, offs, nframes, max_frame_points, cfg, out_points, out_mask, out_mapping, out_npoints, out_coords, counts,
                              w + mine, ws_bytes - mine, st, a.bail);
}

}  // namespace d3d
