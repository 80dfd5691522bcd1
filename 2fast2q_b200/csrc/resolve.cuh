// resolve.cuh — <= m mismatch resolution of packed keys that missed the exact table.
// Rule (fast2q.py:692-750 + 660-690): among library entries of the SAME length take the minimum Hamming
// distance d*; the read is assigned iff d* <= m and exactly one entry attains it.  A symbol of the read that is
// not A/C/G/T ("bad") mismatches every packed library symbol.
//
// Three exact strategies over the same tables:
//   probe  (m == 1)  enumerate the 3*L Hamming-1 neighbours, probe the exact hash table        k_resolve_probe
//   seed   (any m)   pigeonhole: m+1 segments, one must match exactly -> candidates -> XOR+popc  k_resolve_seed
//   scan   (any m)   XOR+popc against library tiles staged in shared memory                      k_resolve_scan
// plus resolve_thread(), the one-thread version used when a queue overflows.
#pragma once

#include "f2q_dev.cuh"

namespace f2q {

constexpr uint32_t RES_NONE = 0xFFFFFFFFu;

// spread the low 32 bits to the even bit positions of a 64-bit word
__host__ __device__ __forceinline__ uint64_t spread_even(uint32_t v) {
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

// number of differing symbols between two 2-bit packed keys, ignoring positions flagged in notbad_even's complement
__host__ __device__ __forceinline__ int packed_distance(uint64_t a, uint64_t b, uint64_t good_even) {
    uint64_t x = a ^ b;
    x = (x | (x >> 1)) & good_even;
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

struct Best {
    int d; uint32_t n, idx;
    __host__ __device__ __forceinline__ void add(int dist, uint32_t i, int m) {
        if (dist < d) { d = dist; n = 1; idx = i; }
        else if (dist == d && dist <= m) n++;
    }
    // the same for candidate lists in which an entry may show up more than once (seeds that overlap): the one entry that
    // is alone at the minimum is not counted twice; once two are there the count only has to stay >= 2
    __host__ __device__ __forceinline__ void add_once(int dist, uint32_t i, int m) {
        if (dist < d) { d = dist; n = 1; idx = i; }
        else if (dist == d && dist <= m && !(n == 1 && idx == i)) n++;
    }
};

#ifdef __CUDACC__
// ---- one thread, any m (overflow fallback; also the reference implementation of the two kernels below) ----
__device__ inline uint32_t resolve_thread(const LibTables& T, int m, uint64_t key, uint32_t bad, uint32_t len) {
    const int nbad = __popc(bad);
    if (m <= 0 || nbad > m) return RES_NONE;
    if (m == 1) {
        uint32_t hits = 0, idx = RES_NONE;
        if (nbad == 1) {
            int p = __ffs(bad) - 1;
            for (uint64_t c = 0; c < 4; c++) {
                uint32_t r = fast_lookup(T, key | (c << (2 * p)), len);
                if (r != SLOT_EMPTY) { hits++; idx = r; }
            }
        } else {
            for (uint32_t p = 0; p < len; p++) {
                for (uint64_t c = 1; c < 4; c++) {
                    uint64_t k2 = key ^ (c << (2 * p));       // (symbol at p) ^ c runs over the three other bases
                    uint32_t r = fast_lookup(T, k2, len);
                    if (r != SLOT_EMPTY) { hits++; idx = r; }
                }
            }
        }
        return hits == 1 ? idx : RES_NONE;
    }
    const uint64_t good = (len >= 32 ? 0x5555555555555555ull : ((1ull << (2 * len)) - 1) & 0x5555555555555555ull) & ~spread_even(bad);
    Best b{m + 1, 0, 0};
    for (uint32_t i = 0; i < T.n_fast; i++) {
        if (__ldg(T.fast_lens + i) != len) continue;
        int d = packed_distance(key, __ldg(T.fast_keys + i), good) + nbad;
        b.add(d, __ldg(T.fast_idx + i), m);
    }
    return (b.d <= m && b.n == 1) ? b.idx : RES_NONE;
}

// ---- Hamming-1 neighbour probing: one warp per queued key, lane p owns position p -----------------
// the queue is cut into n_segs segments of seg_cap entries, one per tile-kernel CTA; seg_count[s] = entries in use
__global__ void __launch_bounds__(256) k_resolve_probe(LibTables T, const QEntry* __restrict__ queue, const uint32_t* __restrict__ seg_count,
                                                       uint32_t seg_cap, uint32_t n_segs, unsigned long long* counts,
                                                       unsigned long long* stats) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t imperfect = 0, nonal = 0;
    for (uint32_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
      const uint32_t n = min(seg_count[seg], seg_cap);
      const QEntry* __restrict__ q = queue + (size_t)seg * seg_cap;
      for (uint32_t i = warp; i < n; i += nwarps) {
        const QEntry e = q[i];
        const int nbad = __popc(e.bad);
        uint32_t hits = 0, idx = 0;
        if (nbad <= 1) {
            if (nbad == 1) {
                const int p = __ffs(e.bad) - 1;
                if (lane < 4) {
                    uint32_t r = fast_lookup(T, e.key | ((uint64_t)lane << (2 * p)), e.len);
                    if (r != SLOT_EMPTY) { hits = 1; idx = r; }
                }
            } else if (lane < e.len) {
                #pragma unroll
                for (uint64_t c = 1; c < 4; c++) {
                    uint32_t r = fast_lookup(T, e.key ^ (c << (2 * lane)), e.len);
                    if (r != SLOT_EMPTY) { hits++; idx = r; }
                }
            }
        }
        const uint32_t total = __reduce_add_sync(0xffffffffu, hits);
        const uint32_t who = __ballot_sync(0xffffffffu, hits != 0);
        if (total == 1) {
            if (lane == (uint32_t)(__ffs(who) - 1)) atomicAdd(counts + idx, 1ull);
            imperfect++;
        } else nonal++;
      }
    }
    if (lane == 0) {
        if (imperfect) atomicAdd(stats + F2Q_STAT_IMPERFECT, (unsigned long long)imperfect);
        if (nonal) atomicAdd(stats + F2Q_STAT_NON_ALIGNED, (unsigned long long)nonal);
    }
}

// ---- library tile scan: XOR + popc of every queued key against shared-memory tiles of the packed library ----
constexpr int SCAN_TILE = 2048;     // library entries per shared-memory tile (16 KB keys + 8 KB lens + 8 KB idx)
constexpr int SCAN_THREADS = 256;

__global__ void __launch_bounds__(SCAN_THREADS) k_resolve_scan(LibTables T, int m, const QEntry* __restrict__ queue,
                                                               const uint32_t* __restrict__ seg_count, uint32_t seg_cap, uint32_t n_segs,
                                                               unsigned long long* counts, unsigned long long* stats) {
    __shared__ __align__(16) uint64_t s_key[SCAN_TILE];
    __shared__ uint32_t s_len[SCAN_TILE];
    __shared__ uint32_t s_idx[SCAN_TILE];
    uint32_t imperfect = 0, nonal = 0;
    for (uint32_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
      const uint32_t n = min(seg_count[seg], seg_cap);
      const QEntry* __restrict__ q = queue + (size_t)seg * seg_cap;
      for (uint32_t base = 0; base < n; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const bool live = i < n;
        QEntry e{0, 0, 0};
        if (live) e = q[i];
        const int nbad = __popc(e.bad);
        const uint64_t good = (e.len >= 32 ? 0x5555555555555555ull : ((1ull << (2 * e.len)) - 1) & 0x5555555555555555ull) & ~spread_even(e.bad);
        Best b{m + 1, 0, 0};
        for (uint32_t t0 = 0; t0 < T.n_fast; t0 += SCAN_TILE) {
            const uint32_t cnt = min((uint32_t)SCAN_TILE, T.n_fast - t0);
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < cnt; k += SCAN_THREADS) {
                s_key[k] = __ldg(T.fast_keys + t0 + k);
                s_len[k] = __ldg(T.fast_lens + t0 + k);
                s_idx[k] = __ldg(T.fast_idx + t0 + k);
            }
            __syncthreads();
            if (live && nbad <= m) {
                for (uint32_t k = 0; k < cnt; k++) {                 // every lane reads the same entry: smem broadcast
                    if (s_len[k] != e.len) continue;
                    int d = packed_distance(e.key, s_key[k], good) + nbad;
                    b.add(d, s_idx[k], m);
                }
            }
        }
        if (live) {
            if (b.d <= m && b.n == 1) { atomicAdd(counts + b.idx, 1ull); imperfect++; }
            else nonal++;
        }
      }
    }
    imperfect = __reduce_add_sync(0xffffffffu, imperfect);
    nonal = __reduce_add_sync(0xffffffffu, nonal);
    if ((threadIdx.x & 31) == 0) {
        if (imperfect) atomicAdd(stats + F2Q_STAT_IMPERFECT, (unsigned long long)imperfect);
        if (nonal) atomicAdd(stats + F2Q_STAT_NON_ALIGNED, (unsigned long long)nonal);
    }
}
// ---- pigeonhole seed index: cut a key into P segments; m mismatches spoil at most m of them, so a key within m mismatches
// of a library entry agrees exactly with it on some choice of P - m segments.  A SEED is such a choice (a bit mask over the
// segments); the index hashes (length, seed number, the chosen segments' symbols) -> range of seed_recs (candidate keys stored
// inline, bucket-contiguous).  P = m + 1 gives the classic one-segment seeds: 3 probes but ~36 candidates per key for 100 000
// guides at m = 2 (a 6-7 symbol segment has 4 096 - 16 384 values); P = 4 gives 6 seeds of two segments (10 symbols, a
// million values): 6 probes and ~0.6 candidates.  f2q_set_library picks P from the library's size (seed_plan below).
constexpr int SEED_MAX_PARTS = 8, SEED_MAX_COMBOS = 32;
__host__ __device__ __forceinline__ uint64_t seed_tag(uint32_t len, uint32_t seg, uint64_t v) {
    return (1ull << 63) | ((uint64_t)len << 40) | ((uint64_t)seg << 32) | v;
}
__host__ __device__ __forceinline__ uint32_t seed_hash(uint64_t tag) {
    uint64_t h = tag * 0x9E3779B97F4A7C15ull;
    h ^= h >> 31; h *= 0xD6E8FEB86659FD93ull;
    return (uint32_t)(h >> 32);
}
__host__ __device__ __forceinline__ uint64_t even_range(uint32_t b0, uint32_t b1) {     // even bits of symbols [b0, b1)
    const uint64_t hi = b1 >= 32 ? ~0ull : ((1ull << (2 * b1)) - 1), lo = b0 >= 32 ? ~0ull : ((1ull << (2 * b0)) - 1);
    return (hi & ~lo) & 0x5555555555555555ull;
}

// bounds (optional): segment boundaries s * len / parts for every (len <= 32, s <= parts) as bytes at [len * SEED_BOUND_STRIDE + s]
// (the resolver kernel keeps them in shared memory: the integer divisions were most of its instructions)
constexpr uint32_t SEED_BOUND_STRIDE = 34;
// seed `combo` of a key on the device: the chosen segments' symbols, lowest segment on top (what seed_value builds on the
// host); *skip when one of them holds a non-ACGT symbol (it can never agree exactly).  The work is per CHOSEN segment, so
// the two one-segment seeds of m = 1 cost a fraction of the six two-segment seeds of m = 2
__device__ __forceinline__ uint64_t seed_of(uint32_t combo, uint32_t parts, uint64_t key, uint64_t badeven, uint32_t len, const uint8_t* bl, bool* skip) {
    uint64_t x = 0;
    bool bad = false;
    for (uint32_t s = 0; s < parts; s++) {
        if (!((combo >> s) & 1u)) continue;
        const uint32_t b0 = bl ? bl[s] : s * len / parts, b1 = bl ? bl[s + 1] : (s + 1) * len / parts;
        const uint32_t w = 2 * (b1 - b0);                              // (<= 32 bits: seed_plan)
        const uint64_t m = (1ull << w) - 1ull;
        bad = bad || ((badeven >> (2 * b0)) & m) != 0;
        x = (x << w) | ((key >> (2 * b0)) & m);
    }
    *skip = bad;
    return x;
}
// the host's form (f2q_set_library), from per-segment values and widths
__host__ __device__ __forceinline__ uint64_t seed_value(uint32_t combo, const uint32_t* v, const uint32_t* w) {
    uint64_t x = 0;
    #pragma unroll
    for (int s = 0; s < SEED_MAX_PARTS; s++) if ((combo >> s) & 1u) x = (x << w[s]) | v[s];
    return x;
}
// bucket of a seed: (first record, count)
__device__ __forceinline__ uint2 seed_bucket(const LibTables& T, uint64_t tag) {
    uint32_t h = seed_hash(tag) & T.seed_mask;
    for (;;) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(T.seed_slots) + h);
        const uint64_t t = ((uint64_t)raw.y << 32) | raw.x;
        if (t == 0) return make_uint2(0, 0);
        if (t == tag) return make_uint2(raw.z, raw.w);
        h = (h + 1) & T.seed_mask;
    }
}

// the classic plan (seed_ncombo == 0): miss + 1 one-segment seeds, any number of segments
__device__ __noinline__ uint32_t resolve_seed_classic(const LibTables& T, int m, uint64_t key, uint32_t bad, uint32_t len) {
    const uint32_t parts = T.seed_parts;
    const uint64_t lenmask = even_range(0, len), badeven = spread_even(bad) & lenmask;
    Best b{m + 1, 0, 0};
    for (uint32_t s = 0; s < parts; s++) {
        const uint32_t b0 = s * len / parts, b1 = (s + 1) * len / parts;
        const uint64_t seg = even_range(b0, b1);
        if (badeven & seg) continue;
        const uint2 bk = seed_bucket(T, seed_tag(len, s, (key >> (2 * b0)) & ((seg | (seg << 1)) >> (2 * b0))));
        for (uint32_t k = 0; k < bk.y; k++) {
            const uint4 it = __ldg(T.seed_recs + bk.x + k);
            const uint64_t x = key ^ (((uint64_t)it.y << 32) | it.x);
            const uint64_t diff = (((x | (x >> 1)) & lenmask) | badeven);
            b.add_once(__popcll(diff), it.z, m);
        }
    }
    return (b.d <= m && b.n == 1) ? b.idx : RES_NONE;
}

__device__ __forceinline__ uint32_t resolve_seed_inline(const LibTables& T, int m, uint64_t key, uint32_t bad, uint32_t len, const uint8_t* bounds = nullptr) {
    const int nbad = __popc(bad);
    if (m <= 0 || nbad > m) return RES_NONE;
    if (T.seed_ncombo == 0) return resolve_seed_classic(T, m, key, bad, len);
    const uint64_t lenmask = even_range(0, len), badeven = spread_even(bad) & lenmask;
    const uint8_t* bl = bounds ? bounds + len * SEED_BOUND_STRIDE : nullptr;
    Best b{m + 1, 0, 0};
    for (uint32_t c = 0; c < T.seed_ncombo; c++) {
        bool skip;
        const uint64_t v = seed_of(T.seed_combo[c], T.seed_parts, key, badeven, len, bl, &skip);
        if (skip) continue;
        const uint2 bk = seed_bucket(T, seed_tag(len, c, v));
        for (uint32_t k = 0; k < bk.y; k++) {
            const uint4 it = __ldg(T.seed_recs + bk.x + k);            // {key lo, key hi, feature index, -}: one load per candidate
            const uint64_t x = key ^ (((uint64_t)it.y << 32) | it.x);
            const uint64_t diff = (((x | (x >> 1)) & lenmask) | badeven);     // even bit 2p: symbol p differs (or is bad)
            b.add_once(__popcll(diff), it.z, m);                       // (an entry that agrees on more segments comes through several seeds)
        }
    }
    return (b.d <= m && b.n == 1) ? b.idx : RES_NONE;
}
// out of line, for the callers in which resolving is the rare case (a full queue inside the streaming kernels): inlined
// there it costs the hot loop registers (measured: streaming kernel 3.97 -> 3.69 ms with this call out of line)
__device__ __noinline__ uint32_t resolve_seed_thread(const LibTables& T, int m, uint64_t key, uint32_t bad, uint32_t len, const uint8_t* bounds = nullptr) {
    return resolve_seed_inline(T, m, key, bad, len, bounds);
}

__global__ void __launch_bounds__(256) k_resolve_seed(LibTables T, int m, const QEntry* __restrict__ queue, const uint32_t* __restrict__ seg_count,
                                                      uint32_t seg_cap, uint32_t n_segs, unsigned long long* counts, unsigned long long* stats) {
    __shared__ uint8_t s_bounds[33 * SEED_BOUND_STRIDE];
    for (uint32_t i = threadIdx.x; i < 33u * SEED_BOUND_STRIDE; i += blockDim.x) {
        const uint32_t len = i / SEED_BOUND_STRIDE, sg = i % SEED_BOUND_STRIDE;
        s_bounds[i] = (uint8_t)(min(sg, T.seed_parts) * len / T.seed_parts);
    }
    __syncthreads();
    uint32_t imperfect = 0, nonal = 0;
    for (uint32_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        const uint32_t n = min(seg_count[seg], seg_cap);
        const QEntry* __restrict__ q = queue + (size_t)seg * seg_cap;
        for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {     // gridDim.y CTAs share a segment
            const QEntry e = q[i];
            const uint32_t r = resolve_seed_inline(T, m, e.key, e.bad, e.len, e.len <= 32 ? s_bounds : nullptr);     // (the kernel's own work: inlined, tables in the constant bank)
            if (r != RES_NONE) { atomicAdd(counts + r, 1ull); imperfect++; } else nonal++;
        }
    }
    imperfect = __reduce_add_sync(0xffffffffu, imperfect);
    nonal = __reduce_add_sync(0xffffffffu, nonal);
    if ((threadIdx.x & 31) == 0) {
        if (imperfect) atomicAdd(stats + F2Q_STAT_IMPERFECT, (unsigned long long)imperfect);
        if (nonal) atomicAdd(stats + F2Q_STAT_NON_ALIGNED, (unsigned long long)nonal);
    }
}

// ---- the same pigeonhole resolution, G lanes per key: they share the bucket probes, split the candidates and merge
// (min distance, how many attain it, which) with shuffles: worth it when buckets are long (one-segment seeds of a big library).  Resolved keys
// without bad symbols go through the memo (LibTables::memo) first.
template <int G>
__device__ __forceinline__ uint32_t resolve_seed_group(const LibTables& T, int m, uint64_t key, uint32_t bad, uint32_t len, bool act,
                                                       const uint8_t* bounds, uint32_t gl, unsigned long long* memo_stats) {
    const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
    const bool memo_ok = act && T.memo && bad == 0;
    uint32_t cached = 0;
    if (memo_ok && gl == 0) cached = memo_lookup(T, klo, khi, len);
    if (G > 1) cached = __shfl_sync(0xffffffffu, cached, (threadIdx.x & 31u) - gl);
    if (memo_ok && gl == 0 && memo_stats) { atomicAdd(memo_stats, 1ull); if (cached) atomicAdd(memo_stats + 1, 1ull); }
    const int nbad = __popc(bad);
    Best b{m + 1, 0, 0};
    if (act && !cached && m > 0 && nbad <= m) {
        const uint64_t lenmask = even_range(0, len), badeven = spread_even(bad) & lenmask;
        const uint8_t* bl = bounds + len * SEED_BOUND_STRIDE;
        for (uint32_t c = 0; c < T.seed_ncombo; c++) {
            bool skip;
            const uint64_t v = seed_of(T.seed_combo[c], T.seed_parts, key, badeven, len, bl, &skip);
            if (skip) continue;
            const uint2 bk = seed_bucket(T, seed_tag(len, c, v));
            for (uint32_t k = gl; k < bk.y; k += (uint32_t)G) {
                const uint4 it = __ldg(T.seed_recs + bk.x + k);
                const uint64_t x = key ^ (((uint64_t)it.y << 32) | it.x);
                const uint64_t diff = (((x | (x >> 1)) & lenmask) | badeven);
                b.add_once(__popcll(diff), it.z, m);
            }
        }
    }
    #pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const int od = __shfl_xor_sync(0xffffffffu, b.d, o);
        const uint32_t on = __shfl_xor_sync(0xffffffffu, b.n, o), oi = __shfl_xor_sync(0xffffffffu, b.idx, o);
        if (od < b.d) { b.d = od; b.n = on; b.idx = oi; }
        else if (od == b.d && !(b.n == 1 && on == 1 && oi == b.idx)) b.n += on;        // (the same entry found by two lanes is one entry)
    }
    uint32_t r = (b.d <= m && b.n == 1) ? b.idx : RES_NONE;
    if (cached) r = cached - 2u;                                        // stored: 1 = not aligned (RES_NONE = 1 - 2), idx + 2
    else if (memo_ok && gl == 0) memo_store(T, klo, khi, len, r + 2u);
    return r;
}

template <int G>
__global__ void __launch_bounds__(256) k_resolve_seed_g(LibTables T, int m, const QEntry* __restrict__ queue, const uint32_t* __restrict__ seg_count,
                                                        uint32_t seg_cap, uint32_t n_segs, unsigned long long* counts, unsigned long long* stats,
                                                        unsigned long long* memo_stats) {
    __shared__ uint8_t s_bounds[33 * SEED_BOUND_STRIDE];
    for (uint32_t i = threadIdx.x; i < 33u * SEED_BOUND_STRIDE; i += blockDim.x) {
        const uint32_t len = i / SEED_BOUND_STRIDE, sg = i % SEED_BOUND_STRIDE;
        s_bounds[i] = (uint8_t)(min(sg, T.seed_parts) * len / T.seed_parts);
    }
    __syncthreads();
    const uint32_t gl = threadIdx.x % G, grp = threadIdx.x / G, groups = blockDim.x / G;
    uint32_t imperfect = 0, nonal = 0;
    for (uint32_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        const uint32_t n = min(seg_count[seg], seg_cap);
        const QEntry* __restrict__ q = queue + (size_t)seg * seg_cap;
        const uint32_t per_round = gridDim.y * groups;                  // (uniform trip count: the shuffles need whole warps)
        for (uint32_t base = 0; base < n; base += per_round) {
            const uint32_t i = base + blockIdx.y * groups + grp;
            const bool act = i < n;
            QEntry e; e.key = 0; e.bad = 0; e.len = 0;
            if (act) e = q[i];
            const bool ok = act && e.len <= 32;
            uint32_t r = resolve_seed_group<G>(T, m, e.key, e.bad, min(e.len, 32u), ok, s_bounds, gl, memo_stats);
            if (act && !ok) r = resolve_seed_thread(T, m, e.key, e.bad, e.len);
            if (act && gl == 0) { if (r != RES_NONE) { atomicAdd(counts + r, 1ull); imperfect++; } else nonal++; }
        }
    }
    imperfect = __reduce_add_sync(0xffffffffu, imperfect);
    nonal = __reduce_add_sync(0xffffffffu, nonal);
    if ((threadIdx.x & 31) == 0) {
        if (imperfect) atomicAdd(stats + F2Q_STAT_IMPERFECT, (unsigned long long)imperfect);
        if (nonal) atomicAdd(stats + F2Q_STAT_NON_ALIGNED, (unsigned long long)nonal);
    }
}
#endif  // __CUDACC__

}  // namespace f2q
