// voxel_tiles.cu -- sparse voxelization as a software pipeline of streaming tiles (the default fast path).
//
// Replaces voxelize_sparse + voxelize_filter (reference d3d/voxel/voxelize.cpp:288-484) for the common
// configurations (no voxel cap, max_points filter NONE or TRIM with max_points <= 8); everything else goes to
// the cluster path (voxel_cluster.cu) or the sort path (voxel.cu).  Same packed outputs, bit for bit.
//
// Why not one cluster per frame (voxel_cluster.cu): that kernel keeps a whole frame in the shared memory of 8 SMs,
// which pins one 1024-thread CTA per SM, all warps of an SM in the same phase, on the 120 SMs that 8-CTA clusters
// reach -- it is bound by exposed latency, not by HBM (ncu: issue slots 52 % busy, DRAM 15 %).  Here every stage is
// an ordinary tile of 256 threads, several CTAs per SM on all 148 SMs, and the stages of different frames run side
// by side in one launch ("tick"):
//
//   split   (1024 points / CTA)  cell keys with the reference's fp32 arithmetic; a point's (key, index) goes to
//                                the queue of the BUCKET its key hashes to (128 buckets for a 120k-point frame);
//                                the tile reserves its share of every queue with one atomic per bucket.  The point's
//                                word is its key: a point that hears nothing back is the only point of its voxel.
//   bucket  (one CTA / bucket)   all points of a voxel meet in one bucket: hash table in shared memory (CAS claim,
//                                atomicMin of the point index, count), the K smallest indices of voxels with more
//                                than K points by a min-cascade (level j keeps the j-th smallest of everything it
//                                is offered and passes the larger value on: the result does not depend on the
//                                arrival order); the points that share a voxel get a new word.
//   scan    (8192 points / CTA)  first-of-voxel and keep bits -> ballots -> one chained scan over all tiles of all
//                                frames (decoupled look-back): voxel ids in order of first appearance, packed rows.
//   write   (1024 points / CTA)  voxel rows and kept point rows streamed out; no barrier, no waiting.
//
// Tick k launches split(chunk k) + bucket(chunk k-1) + scan(chunk k-2) + write(chunk k-3), a chunk being a few
// frames, so the scratch of the frames in flight (queues, words, row tables: ~1.3 MB per frame, plus the frame
// itself) lives in L2 and HBM sees the algorithmic traffic only: 16 B/point in, 32 B/kept point + 28 B/voxel out.
//
// Determinism: every output is a function of the point set (smallest index, count, K smallest indices, prefix
// sums in point order), never of the race order of the queues and tables.
//
// Anything that does not fit (a bucket queue or table overflow: e.g. thousands of points in one voxel) raises a
// device flag; the cluster kernel is launched behind the pipeline on that flag and redoes the batch.
#include "voxel.cuh"
#include <stdlib.h>

namespace d3d {

constexpr int VT_THREADS = 256;
constexpr int VT_WARPS = VT_THREADS / 32;
constexpr int VT_PPT = 4;                         // points per thread (split, write)
constexpr int VT_TILE = VT_THREADS * VT_PPT;      // 1024 points per tile
constexpr int VT_SPT = 32;                        // points per thread (scan): lane u of a warp keeps the warp's row u
constexpr int VT_STILE = VT_THREADS * VT_SPT;     // 8192 points per scan tile
constexpr int VT_STAGES = 4;
constexpr int VT_MAXK = 8;                        // deepest min-cascade (max_points of the TRIM filter)
constexpr int VT_POOL = 128;                      // crowded-voxel records per bucket
constexpr uint32_t VT_NONE = 0xffffffffu;         // empty table slot / queue entry
// point word, category in bits 31:30 -- 0: cell key (<= 30 bits) of a point that is alone in its voxel | 1: nothing to write |
// 2: kept point of a shared voxel + index of the voxel's first point | 3: first point of a shared voxel + point count
constexpr uint32_t VT_W_NONE = 1u << 30, VT_W_JOIN = 2u << 30, VT_W_HEAD = 3u << 30, VT_VAL = (1u << 30) - 1;
constexpr unsigned long long VT_PVAL = (1ull << 62) - 1;   // status word: flag << 62 | kept rows << 31 | voxel rows
constexpr uint32_t VT_F31 = 0x7fffffffu;

struct VtGeom {
    uint32_t lmax, lpad, tpf, spf; // longest frame, padded to whole scan tiles, tiles / scan tiles per frame
    uint32_t lgP, P, lgS, S, qcap; // buckets per frame, table slots per bucket (maximum), queue entries per bucket
    uint32_t CF, nchunks;          // frames per chunk, chunks
};

struct VtArgs {
    const float *pts; int nfeat; const int64_t *offs; int64_t nframes;
    float *out_points; int64_t *out_mask; int64_t *out_mapping; int32_t *out_npoints; int64_t *out_coords; int64_t *counts;
    uint2 *queue; uint32_t *word, *hkey; uint4 *rowinfo; uint32_t *qcount;
    unsigned long long *status; uint32_t *tickets, *bail;
    VtGeom g;
    uint32_t K, cthr; int min_points; int trim, dropall, single_ok;
};

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

// Decoupled look-back over the scan tiles of all frames (one warp).  Tiles take their numbers from a ticket counter, so
// a tile only waits for tiles that are running or done.  flag 1: the tile's own sums; flag 2: inclusive prefix.
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
__device__ __forceinline__ uint32_t vt_home(uint32_t key, uint32_t hshift) { return (key * 0x85EBCA6Bu) >> hshift; }

// what a point's word says (shared by scan and write)
__device__ __forceinline__ void vt_decode(const VtArgs &a, uint32_t w, bool *head, bool *keep)
{
    const uint32_t cat = w >> 30;
    *head = ((cat + 1u) & 3u) < 2u;      // categories 0 and 3
    *keep = cat != 1u && !a.dropall;
}

// ------------------------------------------------------------------------------------------------ split
template <bool NF4>
__device__ __forceinline__ void vt_split(const VtArgs &a, const VcDev &dv, uint32_t lt, uint32_t chunk, unsigned char *dyn)
{
    const VtGeom &g = a.g;
    const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const unsigned ltmask = lanemask_lt();
    const uint32_t fl = lt / g.tpf, t = lt - fl * g.tpf;
    const int64_t f = (int64_t)chunk * g.CF + fl;
    const uint32_t slot = (chunk % VT_STAGES) * g.CF + fl;
    const int64_t b = a.offs[f];
    const uint32_t L = (uint32_t)min((long long)(a.offs[f + 1] - b), (long long)g.lmax);
    const uint32_t t0 = t * VT_TILE;
    if (t0 >= L) return;

    // Shared-memory atomics run at ~0.5 lanes per clock: the tile's histogram over the buckets is kept per warp instead
    // (lanes of a warp that hit the same bucket find each other with match.any, the first of them does a plain
    // read-modify-write of the warp's counter), and the warps' counters are chained afterwards.
    uint32_t *base = reinterpret_cast<uint32_t *>(dyn);                         // [P] first queue position of the tile per bucket
    uint16_t *wh = reinterpret_cast<uint16_t *>(base + g.P), *whw = wh + w * g.P;   // [8][P] per-warp counts, then exclusive offsets
    uint32_t *word = a.word + (size_t)slot * g.lpad;
    uint32_t *qcount = a.qcount + (size_t)slot * g.P;
    uint2 *queue = a.queue + (size_t)slot * g.P * g.qcap;

    for (uint32_t p = tid; p < g.P * (VT_WARPS / 2); p += VT_THREADS) reinterpret_cast<uint32_t *>(wh)[p] = 0u;
    __syncthreads();

    float4 p4[VT_PPT];
    bool in[VT_PPT];
#pragma unroll
    for (int u = 0; u < VT_PPT; u++) {
        const uint32_t i = t0 + u * VT_THREADS + tid;
        in[u] = i < L;
        p4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in[u]) {
            if (NF4) p4[u] = __ldg(reinterpret_cast<const float4 *>(a.pts) + b + i);
            else { const float *q = a.pts + (b + i) * a.nfeat; p4[u] = make_float4(q[0], q[1], q[2], 0.f); }
        }
    }
    uint32_t keys[VT_PPT], rank[VT_PPT];
#pragma unroll
    for (int u = 0; u < VT_PPT; u++) {
        uint32_t key;
        const bool ok = vc_cell<false>(dv, p4[u], &key) && in[u];
        keys[u] = ok ? key : VC_NOKEY;
        if (in[u]) word[t0 + u * VT_THREADS + tid] = (ok && a.single_ok) ? key : VT_W_NONE;   // min_points > 1: a lone point is dropped
        const uint32_t bk = ok ? vt_bucket(key, g.lgP) : VT_NONE;
        const unsigned m = __match_any_sync(0xffffffffu, bk);
        const int leader = __ffs((int)m) - 1;
        uint32_t old = 0;
        if (ok && lane == (unsigned)leader) { old = whw[bk]; whw[bk] = (uint16_t)(old + __popc(m)); }
        rank[u] = __shfl_sync(0xffffffffu, old, leader) + __popc(m & ltmask);
        __syncwarp();
    }
    __syncthreads();
    for (uint32_t p = tid; p < g.P; p += VT_THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < VT_WARPS; k++) { const uint32_t c = wh[k * g.P + p]; wh[k * g.P + p] = (uint16_t)run; run += c; }
        base[p] = run ? atomicAdd(&qcount[p], run) : 0u;
    }
    __syncthreads();
    bool over = false;
#pragma unroll
    for (int u = 0; u < VT_PPT; u++) {
        if (keys[u] != VC_NOKEY) {
            const uint32_t bk = vt_bucket(keys[u], g.lgP);
            const uint32_t pos = base[bk] + whw[bk] + rank[u];
            if (pos < g.qcap) queue[bk * g.qcap + pos] = make_uint2(keys[u], t0 + u * VT_THREADS + tid);
            else over = true;
        }
    }
    if (over) *a.bail = 1u;
}

// ------------------------------------------------------------------------------------------------ bucket
// Shared-memory atomics are the scarce resource here too, so the voxels of a bucket are resolved with plain stores: every entry
// of the bucket's queue writes its queue position into the slot its key hashes to if the slot is empty (any writer wins),
// and after a barrier adopts the slot if the winner carries its key, else moves on to the next slot.  Entries of one key move in
// lock step, so they agree on the slot and on the winner; slots only ever fill.  Two such rounds settle most entries, the
// stragglers finish with compare-and-swap.  Only the ~8 % of entries that join another entry's voxel issue atomics (smallest index,
// count) -- on the winner's words.
constexpr uint32_t VT_E_WIN = 1u << 30, VT_E_JOIN = 2u << 30, VT_E_VAL = (1u << 30) - 1;   // entry state: probing slot | winner: smallest index | joiner: winner
constexpr int VT_ROUNDS = 2;

__device__ __forceinline__ void vt_bucket_role(const VtArgs &a, uint32_t lb, uint32_t chunk, unsigned char *dyn)
{
    const VtGeom &g = a.g;
    const unsigned tid = threadIdx.x;
    const uint32_t fl = lb >> g.lgP, bk = lb & (g.P - 1);
    const uint32_t slot_id = (chunk % VT_STAGES) * g.CF + fl;
    const uint32_t K = a.K, cthr = a.cthr;

    uint2 *qs = reinterpret_cast<uint2 *>(dyn);                    // [qcap] the bucket's queue: key, point index
    uint32_t *st = reinterpret_cast<uint32_t *>(qs + g.qcap);      // [qcap] entry state
    uint32_t *cnt = st + g.qcap;                                   // [qcap] winners: points in the voxel | (record + 1) << 16
    uint32_t *slot = cnt + g.qcap;                                 // [S]
    uint32_t *pool = slot + g.S;                                   // VT_POOL records of VT_MAXK indices
    uint32_t *misc = pool + VT_POOL * VT_MAXK;                     // [1] records in use, [2] failure
    uint32_t *qcount = a.qcount + (size_t)slot_id * g.P;
    const uint2 *q = a.queue + ((size_t)slot_id * g.P + bk) * g.qcap;
    uint32_t *word = a.word + (size_t)slot_id * g.lpad, *hkey = a.hkey + (size_t)slot_id * g.lpad;

    const uint32_t n = min(qcount[bk], g.qcap);
    if (n == 0) return;
    // table size for this bucket: load factor <= 2/3 when the maximum allows it
    uint32_t lgS = 6;
    while ((1u << lgS) < n + (n >> 1) && lgS < g.lgS) lgS++;
    const uint32_t S = 1u << lgS, smask = S - 1, hshift = 32u - lgS;
    for (uint32_t s = tid; s < S; s += VT_THREADS) slot[s] = VT_NONE;
    for (uint32_t e = tid; e < n; e += VT_THREADS) {
        const uint2 x = q[e];
        qs[e] = x; st[e] = vt_home(x.x, hshift); cnt[e] = 1u;
    }
    if (tid == 0) { misc[1] = 0; misc[2] = 0; }
    __syncthreads();
    if (tid == 0) qcount[bk] = 0;   // everybody has read it: the counter is ready for the slot's next frame
    if (n >= S) { if (tid == 0) *a.bail = 1u; return; }   // (cannot settle if every entry is its own voxel; n is CTA-uniform)

#pragma unroll 1
    for (int r = 0; r < VT_ROUNDS; r++) {
        for (uint32_t e = tid; e < n; e += VT_THREADS) {
            const uint32_t s = st[e];
            if (!(s >> 30) && slot[s] == VT_NONE) slot[s] = e;   // racy on purpose: any writer wins
        }
        __syncthreads();
        for (uint32_t e = tid; e < n; e += VT_THREADS) {
            const uint32_t s = st[e];
            if (!(s >> 30)) {
                const uint32_t wv = slot[s];
                if (wv == e) st[e] = VT_E_WIN | qs[e].y;
                else if (qs[wv].x == qs[e].x) st[e] = VT_E_JOIN | wv;
                else st[e] = (s + 1) & smask;
            }
        }
        __syncthreads();   // the next round's stores must not overtake this round's lookups
    }
    // stragglers: claim-or-join with compare-and-swap (no lock step needed: slots only ever fill)
    for (uint32_t e = tid; e < n; e += VT_THREADS) {
        uint32_t s = st[e];
        if (!(s >> 30)) {
            const uint32_t key = qs[e].x;
            for (uint32_t it = 0;; it++) {
                uint32_t wv = slot[s];
                if (wv == VT_NONE) { const uint32_t old = atomicCAS(&slot[s], VT_NONE, e); wv = old == VT_NONE ? e : old; }
                if (wv == e) { st[e] = VT_E_WIN | qs[e].y; break; }
                if (qs[wv].x == key) { st[e] = VT_E_JOIN | wv; break; }
                s = (s + 1) & smask;
                if (it >= S) { misc[2] = 1u; st[e] = VT_E_WIN | qs[e].y; break; }   // cannot happen with n < S
            }
        }
    }
    __syncthreads();
    // joiners: smallest index and point count of the voxel, on the winner's words
    for (uint32_t e = tid; e < n; e += VT_THREADS) {
        const uint32_t s = st[e];
        if ((s >> 30) == 2u) { atomicMin(&st[s & VT_E_VAL], VT_E_WIN | qs[e].y); atomicAdd(&cnt[s & VT_E_VAL], 1u); }
    }
    __syncthreads();
    if (misc[2]) { if (tid == 0) *a.bail = 1u; return; }

    // voxels with more than K points: their K smallest indices by a min-cascade
    if (cthr != VT_NONE) {
        bool anyc = false;
        for (uint32_t e = tid; e < n; e += VT_THREADS) {
            const uint32_t s = st[e];
            const uint32_t wv = (s >> 30) == 2u ? (s & VT_E_VAL) : e;
            if ((cnt[wv] & 0xffffu) > cthr) {
                anyc = true;
                if ((st[wv] & VT_E_VAL) == qs[e].y) {   // the voxel's first point opens the record
                    const uint32_t rec = atomicAdd(&misc[1], 1u);
                    if (rec >= (uint32_t)VT_POOL) misc[2] = 1u;
                    else {
                        cnt[wv] |= (rec + 1u) << 16;   // the only writer of this word in this phase; readers look at the low half
                        for (uint32_t j = 0; j < K; j++) pool[rec * VT_MAXK + j] = VT_NONE;
                    }
                }
            }
        }
        if (__syncthreads_or((int)anyc)) {
            if (misc[2]) { if (tid == 0) *a.bail = 1u; return; }
            for (uint32_t e = tid; e < n; e += VT_THREADS) {
                const uint32_t s = st[e];
                const uint32_t wv = (s >> 30) == 2u ? (s & VT_E_VAL) : e;
                const uint32_t c = cnt[wv];
                if ((c & 0xffffu) > cthr) {
                    uint32_t *rec = pool + ((c >> 16) - 1u) * VT_MAXK;
                    uint32_t v = qs[e].y;
                    for (uint32_t j = 0; j < K; j++) {
                        const uint32_t old = atomicMin(&rec[j], v);
                        v = max(old, v);              // the larger value moves on to the next level
                        if (v == VT_NONE) break;
                    }
                }
            }
            __syncthreads();
        }
    }

    // a new word for every point that shares its voxel (a point that hears nothing is the only point of its voxel)
    for (uint32_t e = tid; e < n; e += VT_THREADS) {
        const uint32_t s = st[e];
        const uint32_t wv = (s >> 30) == 2u ? (s & VT_E_VAL) : e;
        const uint32_t c = cnt[wv], total = c & 0xffffu;
        if (total == 1) continue;
        const uint2 x = qs[e];
        const uint32_t mn = st[wv] & VT_E_VAL;
        uint32_t r;
        if ((long long)total < (long long)a.min_points) r = VT_W_NONE;
        else if (x.y == mn) { r = VT_W_HEAD | total; hkey[x.y] = x.x; }
        else {
            const bool kept = !(total > cthr) || x.y <= pool[((c >> 16) - 1u) * VT_MAXK + K - 1];
            r = kept ? (VT_W_JOIN | mn) : VT_W_NONE;
        }
        word[x.y] = r;
    }
}

// ------------------------------------------------------------------------------------------------ scan
// row table entry: .x first-of-voxel bits of the 32-point row, .y voxel rows before the row, .z keep bits, .w kept rows before the row
__device__ __forceinline__ void vt_scan(const VtArgs &a, uint32_t chunk, unsigned char *dyn)
{
    const VtGeom &g = a.g;
    const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    unsigned long long *wsum = reinterpret_cast<unsigned long long *>(dyn);   // [8] warp sums, then exclusive warp bases
    unsigned long long *s_excl = wsum + VT_WARPS;
    uint32_t *s_misc = reinterpret_cast<uint32_t *>(s_excl + 1);

    if (tid == 0) s_misc[0] = atomicAdd(&a.tickets[chunk], 1u);
    __syncthreads();
    const uint32_t lt = s_misc[0];
    const uint32_t fl = lt / g.spf, t = lt - fl * g.spf;
    const int64_t f = (int64_t)chunk * g.CF + fl;
    const uint32_t slot = (chunk % VT_STAGES) * g.CF + fl;
    const int64_t gt = f * (int64_t)g.spf + t;
    const uint32_t L = (uint32_t)min((long long)(a.offs[f + 1] - a.offs[f]), (long long)g.lmax);
    const uint32_t t0 = t * VT_STILE;
    const uint32_t *word = a.word + (size_t)slot * g.lpad;
    uint4 *ri_f = a.rowinfo + (size_t)slot * (g.lpad / 32);

    // lane u of a warp keeps the bits of the warp's row u
    uint32_t myh = 0, myk = 0;
#pragma unroll
    for (int u = 0; u < VT_SPT; u++) {
        const uint32_t i = t0 + (w * VT_SPT + u) * 32 + lane;
        const uint32_t x = i < L ? word[i] : VT_W_NONE;
        bool head, keep;
        vt_decode(a, x, &head, &keep);
        const uint32_t hb = __ballot_sync(0xffffffffu, head), kb = __ballot_sync(0xffffffffu, keep);
        if (lane == (unsigned)u) { myh = hb; myk = kb; }
    }
    // inclusive prefix over the rows of the warp
    uint32_t ih = __popc(myh), ik = __popc(myk);
#pragma unroll
    for (int d = 1; d < VT_SPT; d <<= 1) {
        const uint32_t xh = __shfl_up_sync(0xffffffffu, ih, d), xk = __shfl_up_sync(0xffffffffu, ik, d);
        if (lane >= (unsigned)d) { ih += xh; ik += xk; }
    }
    if (lane == 31) wsum[w] = ((unsigned long long)ik << 31) | ih;
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
            if (t == 0) { a.counts[2 * f] = (long long)((ex >> 31) & VT_F31); a.counts[2 * f + 1] = (long long)(ex & VT_F31); }
            if (gt == a.nframes * (int64_t)g.spf - 1) {
                const unsigned long long e2 = ex + total;
                a.counts[2 * a.nframes] = (long long)((e2 >> 31) & VT_F31); a.counts[2 * a.nframes + 1] = (long long)(e2 & VT_F31);
            }
        }
    }
    __syncthreads();
    const unsigned long long run0 = *s_excl + wsum[w];
    const uint32_t Gh = ((uint32_t)run0 & VT_F31) + ih - __popc(myh), Gk = ((uint32_t)(run0 >> 31) & VT_F31) + ik - __popc(myk);
    ri_f[t * (VT_STILE / 32) + w * VT_SPT + lane] = make_uint4(myh, Gh, myk, Gk);   // rows past the frame hold zero bits: harmless
}

// ------------------------------------------------------------------------------------------------ write
template <bool NF4>
__device__ __forceinline__ void vt_write(const VtArgs &a, const VcDev &dv, uint32_t lt, uint32_t chunk)
{
    const VtGeom &g = a.g;
    const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    const unsigned ltmask = lanemask_lt();
    const uint32_t fl = lt / g.tpf, t = lt - fl * g.tpf;
    const int64_t f = (int64_t)chunk * g.CF + fl;
    const uint32_t slot = (chunk % VT_STAGES) * g.CF + fl;
    const int64_t b = a.offs[f];
    const uint32_t L = (uint32_t)min((long long)(a.offs[f + 1] - b), (long long)g.lmax);
    const uint32_t t0 = t * VT_TILE;
    if (t0 >= L) return;
    const uint32_t vb = (uint32_t)a.counts[2 * f + 1];   // the frame's first voxel row
    const uint32_t *word = a.word + (size_t)slot * g.lpad, *hkey = a.hkey + (size_t)slot * g.lpad;
    const uint4 *ri_f = a.rowinfo + (size_t)slot * (g.lpad / 32);
    const uint32_t mask_y = (1u << (dv.sh_x - dv.sh_y)) - 1u, mask_z = (1u << dv.sh_y) - 1u;
    constexpr int U = 2;   // rows in flight per warp
#pragma unroll 1
    for (int r0 = 0; r0 < VT_PPT; r0 += U) {
        const uint32_t row0 = (t0 >> 5) + w * VT_PPT + r0;
        if (row0 * 32 >= L) break;
        uint32_t wd[U];
        uint4 ri[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i = (row0 + u) * 32 + lane;
            wd[u] = i < L ? word[i] : VT_W_NONE;
            ri[u] = ri_f[row0 + u];
        }
        // everything the rows need from memory in flight together
        float4 p4[U];
        uint32_t ckey[U];
        uint2 rj[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i = (row0 + u) * 32 + lane;
            bool head, keep;
            vt_decode(a, wd[u], &head, &keep);
            p4[u] = make_float4(0.f, 0.f, 0.f, 0.f); ckey[u] = wd[u]; rj[u] = make_uint2(0u, 0u);
            if (keep && NF4) p4[u] = __ldg(reinterpret_cast<const float4 *>(a.pts) + b + i);
            if (head && (wd[u] >> 30)) ckey[u] = hkey[i];
            if (keep && !head) rj[u] = *reinterpret_cast<const uint2 *>(ri_f + ((wd[u] & VT_VAL) >> 5));   // the id lives where the voxel's first point lives
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i = (row0 + u) * 32 + lane;
            const uint32_t x = wd[u];
            bool head, keep;
            vt_decode(a, x, &head, &keep);
            uint32_t vrow = 0;
            if (head) {
                vrow = ri[u].y + __popc(ri[u].x & ltmask);
                const uint32_t c = (x >> 30) ? (x & VT_VAL) : 1u, key = ckey[u];
                long long *co = reinterpret_cast<long long *>(a.out_coords) + (size_t)vrow * 3;
                __stcs(co + 0, (long long)(key >> dv.sh_x) + dv.cadd[0]);
                __stcs(co + 1, (long long)((key >> dv.sh_y) & mask_y) + dv.cadd[1]);
                __stcs(co + 2, (long long)(key & mask_z) + dv.cadd[2]);
                __stcs(a.out_npoints + vrow, (a.trim && c > a.K) ? (int32_t)a.K : (int32_t)c);
            } else if (keep) {
                const uint32_t m = x & VT_VAL;
                vrow = rj[u].y + __popc(rj[u].x & ((1u << (m & 31u)) - 1u));
            }
            if (keep) {
                const size_t o = (size_t)ri[u].w + __popc(ri[u].z & ltmask);
                if (NF4) __stcs(reinterpret_cast<float4 *>(a.out_points) + o, p4[u]);
                else for (int q = 0; q < a.nfeat; q++) a.out_points[o * a.nfeat + q] = a.pts[(b + i) * a.nfeat + q];
                __stcs(reinterpret_cast<long long *>(a.out_mask) + o, (long long)i);
                __stcs(reinterpret_cast<long long *>(a.out_mapping) + o, (long long)(vrow - vb));
            }
        }
    }
}

// One tick of the pipeline.  The grid is cut into rounds of u.x write + u.y scan + u.z bucket + u.w split blocks so that
// every wave of CTAs mixes the four stages (DRAM writes, a latency chain, shared-memory atomics, DRAM reads) instead
// of running them one after the other; a block whose index passes its stage's count has nothing to do.
struct VtTick { uint32_t chunk, n4, n3, n2, n1, u4, u3, u2, u1; };

template <bool NF4>
__global__ void __launch_bounds__(VT_THREADS, 5) vt_tick_kernel(const VtArgs a, const VcDev dv, const VtTick k)
{
    extern __shared__ __align__(16) unsigned char vt_dyn[];
    const uint32_t U = k.u4 + k.u3 + k.u2 + k.u1;
    const uint32_t m = blockIdx.x / U;
    uint32_t j = blockIdx.x - m * U;
    if (j < k.u4) { const uint32_t i = m * k.u4 + j; if (i < k.n4) vt_write<NF4>(a, dv, i, k.chunk - 3); return; }
    j -= k.u4;
    if (j < k.u3) { if (m * k.u3 + j < k.n3) vt_scan(a, k.chunk - 2, vt_dyn); return; }
    j -= k.u3;
    if (j < k.u2) { const uint32_t i = m * k.u2 + j; if (i < k.n2) vt_bucket_role(a, i, k.chunk - 1, vt_dyn); return; }
    j -= k.u2;
    { const uint32_t i = m * k.u1 + j; if (i < k.n1) vt_split<NF4>(a, dv, i, k.chunk, vt_dyn); }
}

// ------------------------------------------------------------------------------------------------ host side
static uint32_t vt_lg2_ceil(uint64_t x) { uint32_t l = 0; while ((1ull << l) < x) l++; return l; }

static int vt_env_cf()
{
    static int cf = -1;   // read once: the environment is a tuning aid, not part of the call
    if (cf < 0) { const char *e = getenv("D3D_B200_VOX_CF"); int v = e ? atoi(e) : 0; cf = v > 0 ? v : 0; }
    return cf;
}

static int vt_env_roles()
{
    static int r = -1;   // tuning only: bit 0 split, 1 bucket, 2 scan, 3 write (later stages may only be dropped together with everything behind them)
    if (r < 0) { const char *e = getenv("D3D_B200_VOX_ROLES"); r = e ? atoi(e) : 15; }
    return r;
}

static bool vt_geom(int64_t max_frame_points, int64_t nframes, VtGeom *g)
{
    if (max_frame_points < 1) max_frame_points = 1;
    if (max_frame_points > (1ll << 21)) return false;            // 2048 buckets of 1024 points; 29-bit indices in the words
    g->lmax = (uint32_t)max_frame_points;
    g->tpf = (g->lmax + VT_TILE - 1) / VT_TILE;
    g->spf = (g->lmax + VT_STILE - 1) / VT_STILE;
    g->lpad = g->spf * VT_STILE;
    g->lgP = vt_lg2_ceil(((uint64_t)g->lmax + 1023) / 1024);
    g->P = 1u << g->lgP;
    const uint32_t per = (g->lmax + g->P - 1) / g->P;            // points per bucket if every point is kept and the hash is even
    g->lgS = vt_lg2_ceil(2ull * per);
    if (g->lgS < 6) g->lgS = 6;
    if (g->lgS > 11) return false;                                // cannot happen with per <= 1024
    g->S = 1u << g->lgS;
    g->qcap = (per + per / 2 + 256 + 3) & ~3u;                   // the bucket's queue is staged in shared memory; more -> device flag
    int cf = vt_env_cf();
    if (cf <= 0) cf = 8;
    if ((int64_t)cf > nframes) cf = (int)(nframes > 0 ? nframes : 1);
    g->CF = (uint32_t)cf;
    g->nchunks = (uint32_t)((nframes + cf - 1) / cf);
    if ((int64_t)nframes * g->tpf >= (1ll << 31)) return false;
    return true;
}

struct VtLayout { size_t queue, word, hkey, rowinfo, zero, qcount, status, tickets, bail, zero_end, total; };

static VtLayout vt_layout(const VtGeom &g, int64_t nframes)
{
    VtLayout l;
    const size_t slots = (size_t)VT_STAGES * g.CF;
    size_t o = 0;
    l.queue = o;   o += align_up(slots * g.P * g.qcap * sizeof(uint2));
    l.word = o;    o += align_up(slots * g.lpad * 4);
    l.hkey = o;    o += align_up(slots * g.lpad * 4);
    l.rowinfo = o; o += align_up(slots * (g.lpad / 32) * sizeof(uint4));
    l.zero = o;
    l.qcount = o;  o += align_up(slots * g.P * 4);
    l.status = o;  o += align_up((size_t)(nframes > 0 ? nframes : 1) * g.spf * 8);
    l.tickets = o; o += align_up((size_t)g.nchunks * 4);
    l.bail = o;    o += 256;
    l.zero_end = o;
    l.total = o;
    return l;
}

bool vox_tiles_supported(const VoxCfg &cfg, int64_t total, int64_t nframes, int64_t max_frame_points)
{
    VcDev d;
    VtGeom g;
    if (cfg.dense || !vc_make_dev(cfg, &d)) return false;
    if (d.sh_x + (uint32_t)vox_bits_for(cfg.ext[0]) > 30u) return false;                // the point words keep two bits for the category
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
    return align_up(vt_layout(g, nframes).total) + 256 + vox_cluster_ws_bytes(total, nframes, max_frame_points);
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
    a.queue = (uint2 *)(w + lay.queue); a.word = (uint32_t *)(w + lay.word); a.hkey = (uint32_t *)(w + lay.hkey);
    a.rowinfo = (uint4 *)(w + lay.rowinfo); a.qcount = (uint32_t *)(w + lay.qcount);
    a.status = (unsigned long long *)(w + lay.status); a.tickets = (uint32_t *)(w + lay.tickets); a.bail = (uint32_t *)(w + lay.bail);
    a.g = g;
    const bool trim = cfg.pfilter == D3D_PF_TRIM;
    a.K = cfg.max_points > 0 ? (uint32_t)cfg.max_points : 0u;
    a.trim = trim ? 1 : 0;
    a.dropall = (trim && a.K == 0) ? 1 : 0;
    a.cthr = (trim && a.K > 0) ? a.K : VT_NONE;
    a.min_points = cfg.min_points;
    a.single_ok = cfg.min_points <= 1 ? 1 : 0;

    static int smem_set[64];
    const uint32_t dyn_bucket = 16u * g.qcap + 4u * g.S + VT_POOL * VT_MAXK * 4u + 64u;
    const uint32_t dyn_split = 20u * g.P;
    uint32_t dyn = dyn_bucket > dyn_split ? dyn_bucket : dyn_split;
    if (dyn < 256u) dyn = 256u;
    int dev = 0;
    D3D_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !smem_set[dev]) {
        D3D_CUDA_TRY(cudaFuncSetAttribute(vt_tick_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        D3D_CUDA_TRY(cudaFuncSetAttribute(vt_tick_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        smem_set[dev] = 1;
    }
    D3D_CUDA_TRY(cudaMemsetAsync(w + lay.zero, 0, lay.zero_end - lay.zero, st));
    auto nf = [&](int64_t c) -> uint32_t {
        if (c < 0 || c >= (int64_t)g.nchunks) return 0u;
        const int64_t r = nframes - c * g.CF;
        return (uint32_t)(r < (int64_t)g.CF ? r : g.CF);
    };
    for (int64_t k = 0; k < (int64_t)g.nchunks + VT_STAGES - 1; k++) {
        VtTick tk;
        tk.chunk = (uint32_t)k;
        tk.n1 = nf(k) * g.tpf; tk.n2 = nf(k - 1) * g.P; tk.n3 = nf(k - 2) * g.spf; tk.n4 = nf(k - 3) * g.tpf;
        const int roles = vt_env_roles();
        if (!(roles & 1)) tk.n1 = 0;
        if (!(roles & 2)) tk.n2 = 0;
        if (!(roles & 4)) tk.n3 = 0;
        if (!(roles & 8)) tk.n4 = 0;
        const uint32_t total_blocks = tk.n1 + tk.n2 + tk.n3 + tk.n4;
        if (total_blocks == 0) continue;
        const uint32_t rounds = (total_blocks + 47) / 48;   // ~48 blocks per round
        tk.u1 = (tk.n1 + rounds - 1) / rounds; tk.u2 = (tk.n2 + rounds - 1) / rounds;
        tk.u3 = (tk.n3 + rounds - 1) / rounds; tk.u4 = (tk.n4 + rounds - 1) / rounds;
        const uint32_t grid = rounds * (tk.u1 + tk.u2 + tk.u3 + tk.u4);
        if (nfeat == 4) vt_tick_kernel<true><<<grid, VT_THREADS, dyn, st>>>(a, dv, tk);
        else vt_tick_kernel<false><<<grid, VT_THREADS, dyn, st>>>(a, dv, tk);
        D3D_LAUNCHED();
    }
    // the batch is redone by the cluster kernel when a queue, a table or a record pool overflowed (device flag)
    return vox_cluster_sparse(points, total, nfeat, offs, nframes, max_frame_points, cfg, out_points, out_mask, out_mapping, out_npoints, out_coords, counts,
                              w + mine, ws_bytes - mine, st, a.bail);
}

}  // namespace d3d
#ifdef VT_ROLE_PROBE
namespace d3d {
__global__ void __launch_bounds__(VT_THREADS) vt_probe_split(const VtArgs a, const VcDev dv) { extern __shared__ __align__(16) unsigned char d[]; vt_split<true>(a, dv, blockIdx.x, 0, d); }
__global__ void __launch_bounds__(VT_THREADS) vt_probe_bucket(const VtArgs a, const VcDev dv) { extern __shared__ __align__(16) unsigned char d[]; vt_bucket_role(a, blockIdx.x, 0, d); }
__global__ void __launch_bounds__(VT_THREADS) vt_probe_scan(const VtArgs a, const VcDev dv) { extern __shared__ __align__(16) unsigned char d[]; vt_scan(a, 0, d); }
__global__ void __launch_bounds__(VT_THREADS) vt_probe_write(const VtArgs a, const VcDev dv) { vt_write<true>(a, dv, blockIdx.x, 0); }
}
#endif
