// tuning probe (not part of the product): cost of building one hash table per frame in L2 with ONE 64-bit atomicMin per probe
// (the larger of {resident, incoming} entry moves on: "displacing min"), and of the lookup pass behind it, on C2-shaped frames.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o l2hash_probe l2hash_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <unordered_map>
#include <cuda_runtime.h>

constexpr unsigned long long EMPTY = ~0ull;
struct Geo { float size[3]; int vlo[3]; uint32_t ext[3]; uint32_t sh_x, sh_y; };

__device__ __forceinline__ bool cell(const Geo &c, const float4 &p, uint32_t *key)
{
    const float v0 = __fdiv_rn(p.x, c.size[0]), v1 = __fdiv_rn(p.y, c.size[1]), v2 = __fdiv_rn(p.z, c.size[2]);
    const int i0 = __float2int_rd(v0), i1 = __float2int_rd(v1), i2 = __float2int_rd(v2);
    const uint32_t c0 = (uint32_t)(i0 - c.vlo[0]), c1 = (uint32_t)(i1 - c.vlo[1]), c2 = (uint32_t)(i2 - c.vlo[2]);
    *key = (c0 << c.sh_x) | (c1 << c.sh_y) | c2;
    return v0 == v0 && v1 == v1 && v2 == v2 && c0 < c.ext[0] && c1 < c.ext[1] && c2 < c.ext[2];
}
__device__ __forceinline__ uint32_t home(uint32_t key, uint32_t nslots) { return __umulhi(key * 0x9E3779B1u, nslots); }

template <int CTAS>
__global__ void __launch_bounds__(256, CTAS) k_build(const float4 *pts, uint32_t L, unsigned long long *tab, uint32_t nslots, uint32_t *word, Geo g)
{
    const uint32_t f = blockIdx.y, t0 = blockIdx.x * 1024u, tid = threadIdx.x;
    const float4 *p = pts + (size_t)f * L;
    unsigned long long *T = tab + (size_t)f * nslots;
    uint32_t *w = word + (size_t)f * L;
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const uint32_t i = t0 + u * 256 + tid; q[u] = i < L ? __ldg(p + i) : make_float4(1e9f, 0, 0, 0); }
    unsigned long long x[4]; uint32_t s[4]; bool act[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const uint32_t i = t0 + u * 256 + tid;
        uint32_t key;
        act[u] = cell(g, q[u], &key) && i < L;
        if (i < L) w[i] = act[u] ? key : 0x40000000u;
        x[u] = ((unsigned long long)key << 32) | i;
        s[u] = home(key, nslots);
    }
    bool any = act[0] || act[1] || act[2] || act[3];
    while (any) {
        unsigned long long old[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (act[u]) old[u] = atomicMin(T + s[u], x[u]);
        any = false;
#pragma unroll
        for (int u = 0; u < 4; u++) if (act[u]) {
            if (old[u] == EMPTY || (uint32_t)(old[u] >> 32) == (uint32_t)(x[u] >> 32)) act[u] = false;
            else { x[u] = old[u] > x[u] ? old[u] : x[u]; s[u] = s[u] + 1 == nslots ? 0 : s[u] + 1; any = true; }
        }
    }
}

template <int CTAS>
__global__ void __launch_bounds__(256, CTAS) k_lookup(uint32_t L, const unsigned long long *tab, uint32_t nslots, uint32_t *word, uint32_t *cnt, unsigned long long *stats)
{
    const uint32_t f = blockIdx.y, t0 = blockIdx.x * 1024u, tid = threadIdx.x;
    const unsigned long long *T = tab + (size_t)f * nslots;
    uint32_t *w = word + (size_t)f * L, *c = cnt + (size_t)f * L;
    uint32_t key[4], s[4]; bool act[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const uint32_t i = t0 + u * 256 + tid;
        key[u] = i < L ? w[i] : 0x40000000u;
        act[u] = (key[u] >> 30) == 0 || (key[u] >> 30) == 3;   // a joiner may already have flagged this head's word
        key[u] &= 0x3fffffffu;
        s[u] = home(key[u], nslots);
    }
    uint32_t heads = 0, joins = 0, probes = 0;
    uint32_t h[4];
    bool any = act[0] || act[1] || act[2] || act[3];
    bool pend[4] = {act[0], act[1], act[2], act[3]};
    while (any) {
        unsigned long long v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (pend[u]) { v[u] = __ldcg(T + s[u]); probes++; }
        any = false;
#pragma unroll
        for (int u = 0; u < 4; u++) if (pend[u]) {
            if ((uint32_t)(v[u] >> 32) == key[u]) { pend[u] = false; h[u] = (uint32_t)v[u]; }
            else { s[u] = s[u] + 1 == nslots ? 0 : s[u] + 1; any = true; }
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) if (act[u]) {
        const uint32_t i = t0 + u * 256 + tid;
        if (h[u] == i) heads++;
        else { joins++; w[i] = 0x80000000u | h[u]; atomicAdd(c + h[u], 1u); atomicOr(w + h[u], 0xc0000000u); }
    }
    // statistics (probe only)
    for (int d = 16; d; d >>= 1) { heads += __shfl_xor_sync(~0u, heads, d); joins += __shfl_xor_sync(~0u, joins, d); probes += __shfl_xor_sync(~0u, probes, d); }
    if ((tid & 31) == 0) { atomicAdd(stats + 0, (unsigned long long)heads); atomicAdd(stats + 1, (unsigned long long)joins); atomicAdd(stats + 2, (unsigned long long)probes); }
}


// ---- second experiment: two bitmaps per frame (seen once / seen twice) as an exact singleton filter in front of the bucket path
__global__ void __launch_bounds__(256, 8) k_mark(const float4 *pts, uint32_t L, uint32_t *seen1, uint32_t *seen2, uint32_t lgbits, uint32_t *word, Geo g)
{
    const uint32_t f = blockIdx.y, t0 = blockIdx.x * 1024u, tid = threadIdx.x;
    const float4 *p = pts + (size_t)f * L;
    uint32_t *s1 = seen1 + ((size_t)f << (lgbits - 5)), *s2 = seen2 + ((size_t)f << (lgbits - 5));
    uint32_t *w = word + (size_t)f * L;
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const uint32_t i = t0 + u * 256 + tid; q[u] = i < L ? __ldg(p + i) : make_float4(1e9f, 0, 0, 0); }
    uint32_t old[4], bit[4], wd[4]; bool act[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const uint32_t i = t0 + u * 256 + tid;
        uint32_t key;
        act[u] = cell(g, q[u], &key) && i < L;
        if (i < L) w[i] = act[u] ? key : 0x40000000u;
        const uint32_t h = (key * 0x9E3779B1u) >> (32 - lgbits);
        wd[u] = h >> 5; bit[u] = 1u << (h & 31);
        if (act[u]) old[u] = atomicOr(s1 + wd[u], bit[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) if (act[u] && (old[u] & bit[u])) atomicOr(s2 + wd[u], bit[u]);
}
__global__ void __launch_bounds__(256, 8) k_filter(uint32_t L, const uint32_t *seen2, uint32_t lgbits, const uint32_t *word, unsigned long long *stats)
{
    const uint32_t f = blockIdx.y, t0 = blockIdx.x * 1024u, tid = threadIdx.x;
    const uint32_t *s2 = seen2 + ((size_t)f << (lgbits - 5));
    const uint32_t *w = word + (size_t)f * L;
    uint32_t shared_pts = 0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const uint32_t i = t0 + u * 256 + tid;
        const uint32_t key = i < L ? w[i] : 0x40000000u;
        if ((key >> 30) == 0) {
            const uint32_t h = (key * 0x9E3779B1u) >> (32 - lgbits);
            shared_pts += (__ldcg(s2 + (h >> 5)) >> (h & 31)) & 1u;
        }
    }
    __shared__ uint32_t tot;
    if (tid == 0) tot = 0;
    __syncthreads();
    for (int d = 16; d; d >>= 1) shared_pts += __shfl_xor_sync(~0u, shared_pts, d);
    if ((tid & 31) == 0 && shared_pts) atomicAdd(&tot, shared_pts);
    __syncthreads();
    if (tid == 0) atomicAdd(stats + (blockIdx.x & 7), (unsigned long long)tot);
}

static float frand(uint64_t &s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (float)((s >> 40) & 0xffffff) / 16777216.0f; }

int main(int argc, char **argv)
{
    const int NF = argc > 1 ? atoi(argv[1]) : 64;
    const uint32_t L = 120000;
    Geo g;
    g.size[0] = 0.05f; g.size[1] = 0.05f; g.size[2] = 0.1f;
    g.vlo[0] = 0; g.vlo[1] = -800; g.vlo[2] = -30; g.ext[0] = 1408; g.ext[1] = 1600; g.ext[2] = 40;
    g.sh_y = 6; g.sh_x = 6 + 11;
    std::vector<float4> h((size_t)NF * L);
    uint64_t seed = 12345;
    for (size_t i = 0; i < h.size(); i++) {
        const float u = frand(seed), rho = 80.f * u * u, th = (frand(seed) - .5f) * 3.6f;
        float z = 0; for (int k = 0; k < 6; k++) z += frand(seed); z = (z - 3.f) * 0.85f * 0.6f - 1.2f;
        h[i] = make_float4(rho * cosf(th), rho * sinf(th), z, frand(seed));
    }
    // host truth: voxels / joiners over all frames
    unsigned long long tv = 0, tj = 0, crowded = 0, crowded_pts = 0;
    for (int f = 0; f < NF; f++) {
        std::unordered_map<uint32_t, uint32_t> m;
        for (uint32_t i = 0; i < L; i++) {
            const float4 p = h[(size_t)f * L + i];
            const float v0 = p.x / g.size[0], v1 = p.y / g.size[1], v2 = p.z / g.size[2];
            const uint32_t c0 = (uint32_t)((int)floorf(v0) - g.vlo[0]), c1 = (uint32_t)((int)floorf(v1) - g.vlo[1]), c2 = (uint32_t)((int)floorf(v2) - g.vlo[2]);
            if (c0 < g.ext[0] && c1 < g.ext[1] && c2 < g.ext[2]) m[(c0 << g.sh_x) | (c1 << g.sh_y) | c2]++;
        }
        tv += m.size();
        for (auto &kv : m) { tj += kv.second - 1; if (kv.second > 5) { crowded++; crowded_pts += kv.second; } }
    }
    printf("host: %d frames, voxels %llu, joiners %llu, voxels with > 5 points %llu (%llu points)\n", NF, tv, tj, crowded, crowded_pts);

    float4 *pts; uint32_t *word, *cnt; unsigned long long *tab, *stats;
    cudaMalloc(&pts, h.size() * 16); cudaMemcpy(pts, h.data(), h.size() * 16, cudaMemcpyHostToDevice);
    cudaMalloc(&word, (size_t)NF * L * 4); cudaMalloc(&cnt, (size_t)NF * L * 4); cudaMalloc(&stats, 64);
    const uint32_t maxslots = L * 3;
    cudaMalloc(&tab, (size_t)NF * maxslots * 8);
    cudaEvent_t e[4]; for (auto &x : e) cudaEventCreate(&x);
    const float factors[] = {1.25f, 1.5f, 2.2f, 3.0f};
    for (int ctas = 6; ctas <= 8 && argc <= 2; ctas += 2)
    for (float fac : factors) {
        const uint32_t nslots = (uint32_t)(L * fac);
        float best[3] = {1e9f, 1e9f, 1e9f};
        unsigned long long st[3];
        for (int rep = 0; rep < 5; rep++) {
            cudaMemsetAsync(stats, 0, 64);
            cudaEventRecord(e[0]);
            cudaMemsetAsync(tab, 0xff, (size_t)NF * nslots * 8);
            cudaMemsetAsync(cnt, 0, (size_t)NF * L * 4);
            cudaEventRecord(e[1]);
            if (ctas == 8) k_build<8><<<dim3((L + 1023) / 1024, NF), 256>>>(pts, L, tab, nslots, word, g);
            else k_build<6><<<dim3((L + 1023) / 1024, NF), 256>>>(pts, L, tab, nslots, word, g);
            cudaEventRecord(e[2]);
            if (ctas == 8) k_lookup<8><<<dim3((L + 1023) / 1024, NF), 256>>>(L, tab, nslots, word, cnt, stats);
            else k_lookup<6><<<dim3((L + 1023) / 1024, NF), 256>>>(L, tab, nslots, word, cnt, stats);
            cudaEventRecord(e[3]);
            cudaEventSynchronize(e[3]);
            for (int k = 0; k < 3; k++) { float ms; cudaEventElapsedTime(&ms, e[k], e[k + 1]); if (ms < best[k]) best[k] = ms; }
            cudaMemcpy(st, stats, 24, cudaMemcpyDeviceToHost);
        }
        printf("ctas/SM %d slots %.2f L (%.1f MB per %d frames): memset %.1f us, build %.1f us, lookup %.1f us | heads %llu joiners %llu lookup probes %llu  %s  [%s]\n",
               ctas, fac, (double)NF * nslots * 8 / 1e6, NF, best[0] * 1e3, best[1] * 1e3, best[2] * 1e3, st[0], st[1], st[2],
               (st[0] == tv && st[1] == tj) ? "MATCH" : "MISMATCH", cudaGetErrorString(cudaGetLastError()));
    }
    {
        uint32_t *s1, *s2;
        cudaMalloc(&s1, (size_t)NF << 19); cudaMalloc(&s2, (size_t)NF << 19);
        for (uint32_t lgbits = 19; lgbits <= 22; lgbits++) {
            float best[3] = {1e9f, 1e9f, 1e9f};
            unsigned long long st[8];
            for (int rep = 0; rep < 5; rep++) {
                cudaMemsetAsync(stats, 0, 64);
                cudaEventRecord(e[0]);
                cudaMemsetAsync(s1, 0, (size_t)NF << (lgbits - 3)); cudaMemsetAsync(s2, 0, (size_t)NF << (lgbits - 3));
                cudaEventRecord(e[1]);
                k_mark<<<dim3((L + 1023) / 1024, NF), 256>>>(pts, L, s1, s2, lgbits, word, g);
                cudaEventRecord(e[2]);
                k_filter<<<dim3((L + 1023) / 1024, NF), 256>>>(L, s2, lgbits, word, stats);
                cudaEventRecord(e[3]);
                cudaEventSynchronize(e[3]);
                for (int k = 0; k < 3; k++) { float ms; cudaEventElapsedTime(&ms, e[k], e[k + 1]); if (ms < best[k]) best[k] = ms; }
                cudaMemcpy(st, stats, 64, cudaMemcpyDeviceToHost);
            }
            unsigned long long sh = 0; for (int k = 0; k < 8; k++) sh += st[k];
            printf("bitmaps of 2^%u bits per frame: memset %.1f us, mark %.1f us, filter %.1f us | points sent to the bucket path %llu of %llu in range (%.1f %%; truly shared %.1f %%)  [%s]\n",
                   lgbits, best[0] * 1e3, best[1] * 1e3, best[2] * 1e3, sh, tv + tj, 100.0 * sh / (tv + tj), 100.0 * (tj + (tv - (tv + tj - 2 * tj > 0 ? 0 : 0))) / (tv + tj), cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
