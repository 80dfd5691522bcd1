// flex.cuh — device side of the bit-parallel path (flex_core.h) for one read inside the streaming kernel (spec.cuh):
//   extraction   flex_pieces: search sequences with mismatches / several windows per read, from bit planes
//   key          <= 2 pieces of <= 32 ACGT symbols, <= 40 symbols in all, packed 2 bits per symbol into 80 bits + a signature
//                (piece count and lengths), so that 'X:Y' keys (fast2q.py:362-363) are ONE table probe like single keys
//   Counter      exact probe of the flex table (16-byte slots); non-exact keys go to this CTA's queue segment (FlexQ) and are
//                resolved by k_resolve_flex (resolve.cuh) — reads the packed form cannot express go to the generic queue
//   Extract+Count  the key goes to this CTA's segment of the INSERT LOG (8 bytes per read); k_ec_commit inserts the log
//                into the packed key table only when the chunk's speculation verified (spec.cuh), so a failed guess
//                leaves no trace in the table.  Keys the packed table cannot hold (other symbols than ACGT, > 29 symbols,
//                several pieces) go to the generic queue and from there into the byte-arena table.
#pragma once

#include "f2q_dev.cuh"
#include "flex_core.h"
#include "generic.cuh"
#include "tile.cuh"

namespace f2q {

// ---- packed key of the flex path ------------------------------------------------------------------------------
constexpr uint32_t FX_MAX_SYMBOLS = 40;
constexpr uint32_t FX_MAX_BYTELEN = 2 * FLEX_MAX_PIECE + 1;              // longest key string that could be packed
constexpr uint16_t FX_SIG_MIXED = 0xFFFFu;                               // keys of one byte length with different signatures
constexpr uint32_t FX_EMPTY = 0xFFFFFFFFu;

struct FlexKey {
    uint64_t lo;             // symbols 0 .. 31
    uint32_t hi;             // symbols 32 .. 39 (16 bits)
    uint32_t sig;            // len0 | len1 << 6 | pieces << 12
    uint64_t bad;            // symbols whose upper-cased byte is not A/C/G/T (their code bits are 0)
};
__host__ __device__ __forceinline__ uint32_t fx_sig(uint32_t np, uint32_t l0, uint32_t l1) { return l0 | (l1 << 6) | (np << 12); }
__host__ __device__ __forceinline__ uint32_t fx_sig_symbols(uint32_t sig) { return (sig & 63u) + ((sig >> 6) & 63u); }
__host__ __device__ __forceinline__ uint32_t fx_sig_bytelen(uint32_t sig) { const uint32_t np = sig >> 12; return fx_sig_symbols(sig) + (np ? np - 1u : 0u); }

__host__ __device__ __forceinline__ uint32_t fx_hash(uint64_t lo, uint32_t hs) {
    uint64_t h = (lo ^ ((uint64_t)hs * 0x9E3779B97F4A7C15ull)) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
    h *= 0x94D049BB133111EBull;
    return (uint32_t)(h >> 32);
}

// non-exact key waiting for the resolver (one per read, 32 bytes)
struct __align__(16) FlexQ {
    uint64_t lo, bad;
    uint32_t hi, sig, pad0, pad1;
};

// pieces -> key; false when the pieces do not fit (more than FX_MAX_SYMBOLS symbols)
__host__ __device__ __forceinline__ bool fx_assemble(const FlexPiece* pc, int np, FlexKey& k) {
    const uint32_t l0 = pc[0].len, l1 = np > 1 ? pc[1].len : 0u;
    if (l0 + l1 > FX_MAX_SYMBOLS) return false;
    k.lo = pc[0].codes; k.hi = 0; k.bad = pc[0].notok;
    if (np > 1) {
        const uint64_t c1 = pc[1].codes;
        if (l0 < 32u) {
            k.lo |= c1 << (2u * l0);
            k.hi = l0 ? (uint32_t)(c1 >> (64u - 2u * l0)) : 0u;
        } else k.hi = (uint32_t)c1;
        k.bad |= (uint64_t)pc[1].notok << l0;
    }
    k.sig = fx_sig((uint32_t)np, l0, l1);
    return true;
}

#ifdef __CUDACC__
// exact probe of the flex table: value (feature index) or FX_EMPTY.  Slot = {lo.x, lo.y, hi | sig << 16, value}
__device__ __forceinline__ uint32_t fx_lookup(const LibTables& T, const FlexKey& k) {
    const uint32_t z = k.hi | (k.sig << 16);
    uint32_t h = fx_hash(k.lo, z) & T.fx_mask;
    #pragma unroll 1
    for (;;) {
        const uint4 s = __ldg(T.fx_slots + h);
        if (s.w == FX_EMPTY) return FX_EMPTY;
        if (s.x == (uint32_t)k.lo && s.y == (uint32_t)(k.lo >> 32) && s.z == z) return s.w;
        h = (h + 1u) & T.fx_mask;
    }
}

// symbols flagged notok by the planes (raw byte not 'A','C','G','T'): lower-case bases are good symbols of the KEY
// (the piece is upper-cased, fast2q.py:355); their codes are filled in from the bytes.  Rare, byte-wise.
__device__ __noinline__ FlexPiece fx_fix_case(const uint8_t* seq, FlexPiece p) {
    uint32_t todo = p.notok;
    while (todo) {
        const uint32_t i = (uint32_t)__ffs((int)todo) - 1u;
        todo &= todo - 1u;
        uint32_t code;
        if (base_code(seq[p.off + i], code)) { p.notok &= ~(1u << i); p.codes |= (uint64_t)code << (2u * i); }
    }
    return p;
}

// Extract+Count insert log entry: (codes << 6 | len) + 1, len <= 29
constexpr uint32_t EC_PK_MAX_LEN = 29;
__host__ __device__ __forceinline__ unsigned long long ec_pk_tag(uint64_t codes, uint32_t len) { return ((codes << 6) | len) + 1ull; }

// what the streaming kernel needs for the flex policies.  Queue (Counter: FlexQ) and insert log (Extract+Count: u64) are ONE
// array for the whole grid; a warp takes blocks of FX_BLOCK entries from a global counter (DevState::q_count) and fills the
// rest of its last block with empty entries, so that the consumers see [0, min(q_count, capacity)) fully written.  How the
// reads spread over the CTAs (a small chunk is parsed by a few of them) therefore never matters.
constexpr uint32_t FX_BLOCK = 128, FX_NOBLOCK = 0xFFFFFFFFu;
struct FlexCtx {
    FlexQ* q;                          // Counter: the queue
    unsigned long long* log;           // Extract+Count: the insert log
    uint32_t q_cap;                    // entries (multiple of FX_BLOCK)
    uint32_t* hist;                    // shared-memory histogram or nullptr
    GEntry* gqueue;
    DevState* St;
    int mode, miss;
};
struct FlexWarp { uint32_t base, used; };                               // the warp's current block (uniform over its lanes)

__device__ __forceinline__ void fx_block_fill(const FlexCtx& F, const FlexWarp& W, uint32_t lane) {
    if (W.base == FX_NOBLOCK) return;
    for (uint32_t i = W.used + lane; i < FX_BLOCK; i += 32u) {
        if (F.mode == F2Q_MODE_COUNT) { FlexQ e; e.lo = 0; e.bad = 0; e.hi = 0; e.sig = 0; e.pad0 = 0; e.pad1 = 0; F.q[W.base + i] = e; }
        else F.log[W.base + i] = 0ull;
    }
}

// one read of a CONVERGED warp (every lane calls it; `valid` lanes hold a read).  tile = the stage in shared memory,
// [s0, e0) / [s3, e3) the sequence / quality line before rstrip, gseq / gqual their global addresses.  FC: the kernel's own
// parameter copy of the configuration (constant bank: the search-sequence loops run on uniform registers).
template <int PW, int K>
__device__ __forceinline__ void flex_read_warp(const FlexCtx& F, const FlexCfg& FC, bool valid, const uint8_t* tile, uint32_t s0, uint32_t e0,
                                               uint32_t s3, uint32_t e3, const uint8_t* gseq, const uint8_t* gqual, const LibTables& T,
                                               const Outputs& O, Fast1Counts& n, FlexWarp& W, uint32_t lane) {
    if (!valid) { s0 = e0 = s3 = e3 = 0; }
    else {
        if (is_py_space(tile[e0 - 1])) while (e0 > s0 && is_py_space(tile[e0 - 1])) e0--;
        if (is_py_space(tile[e3 - 1])) while (e3 > s3 && is_py_space(tile[e3 - 1])) e3--;
    }
    const uint32_t r = e0 - s0, q = e3 - s3;
    bool generic = valid && (r > 32u * PW || q > 32u * PW);           // longer than the planes: byte-wise code
    const bool live = valid && !generic;
    FlexPiece pc[FLEX_ITER];
    pc[0].codes = 0; pc[0].notok = 0; pc[0].len = 0; pc[0].off = 0;
    pc[1] = pc[0];
    // (the planes are built only for the 8-byte groups the longest line of the warp reaches)
    const uint32_t maxlen = __reduce_max_sync(0xffffffffu, live ? max(r, q) : 0u);
    int np = flex_pieces<PW, K>(FC, tile, live ? s0 : 0u, live ? r : 0u, tile, live ? s3 : 0u, live ? q : 0u, maxlen, pc);
    if (!live) np = -1;
    n.qfail += (live && np == -1) ? 1u : 0u;
    if (live && np <= -2) {
        // a piece longer than 32 symbols.  Counter: nothing can align when the library has no key of the key's byte length
        // (the usual case: a search sequence that also occurs by chance elsewhere in the read); otherwise the generic code
        const uint32_t pieces = (uint32_t)(-np - 1);
        const uint32_t bl = pc[0].len + (pieces > 1 ? pc[1].len + 1u : 0u);
        if (F.mode == F2Q_MODE_COUNT && (bl > FX_MAX_BYTELEN || __ldg(T.fx_len_sig + bl) == 0)) n.nonal++;
        else generic = true;
    }
    bool keyed = live && np >= 1;
    if (keyed && ((pc[0].notok | (np > 1 ? pc[1].notok : 0u)) != 0u)) {                  // rare: lower case / N inside the key
        pc[0] = fx_fix_case(tile + s0, pc[0]);
        if (np > 1) pc[1] = fx_fix_case(tile + s0, pc[1]);
    }
    FlexKey k;
    k.lo = 0; k.hi = 0; k.sig = 0; k.bad = 0;
    if (keyed && !fx_assemble(pc, np, k)) {
        // more than 40 symbols in all: Counter mode needs the generic code only when the library has keys of that byte length
        keyed = false;
        const uint32_t bl = pc[0].len + (np > 1 ? pc[1].len + 1u : 0u);
        if (F.mode == F2Q_MODE_COUNT && (bl > FX_MAX_BYTELEN || __ldg(T.fx_len_sig + bl) == 0)) n.nonal++;
        else generic = true;
    }
    bool to_queue = false;
    if (F.mode == F2Q_MODE_COUNT) {
        uint32_t v = FX_EMPTY;
        if (keyed && k.bad == 0) v = fx_lookup(T, k);
        const bool hit = keyed && v != FX_EMPTY;
        n.perfect += hit ? 1u : 0u;
        if (hit) { if (F.hist) atomicAdd(F.hist + v, 1u); else atomicAdd(O.counts + v, 1ull); }
        if (keyed && !hit) {
            // no exact entry.  Candidates of the mismatch search have the key's BYTE length (fast2q.py:670-672)
            const uint32_t bl = fx_sig_bytelen(k.sig);
            const uint32_t cls = bl <= FX_MAX_BYTELEN ? (uint32_t)__ldg(T.fx_len_sig + bl) : 0u;
            if (cls == 0u) n.nonal++;                                  // no library key of that length: nothing can align
            else if (cls != k.sig) generic = true;                     // same length, other shape (or mixed shapes): byte-wise code decides
            else if (F.miss <= 0) n.nonal++;
            else to_queue = true;
        }
    } else {
        // Extract+Count: every key is counted (fast2q.py:382-387)
        if (keyed) {
            if (np == 1 && k.bad == 0 && pc[0].len <= EC_PK_MAX_LEN) to_queue = true;
            else generic = true;
        }
    }
    // queue / log slots of the warp
    const uint32_t mq = __ballot_sync(0xffffffffu, to_queue);
    if (mq) {
        const uint32_t cnt = (uint32_t)__popc(mq);
        if (W.used + cnt > FX_BLOCK) {                                 // (uniform) next block: one global atomic per FX_BLOCK entries
            fx_block_fill(F, W, lane);
            uint32_t nb = 0;
            if (lane == 0) nb = atomicAdd(&F.St->q_count, FX_BLOCK);
            nb = __shfl_sync(0xffffffffu, nb, 0);
            W.base = (nb <= F.q_cap - FX_BLOCK && F.q_cap >= FX_BLOCK) ? nb : FX_NOBLOCK;
            W.used = 0;
        }
        if (to_queue) {
            if (W.base != FX_NOBLOCK) {
                const uint32_t sl = W.base + W.used + (uint32_t)__popc(mq & ((1u << lane) - 1u));
                if (F.mode == F2Q_MODE_COUNT) { FlexQ e; e.lo = k.lo; e.bad = k.bad; e.hi = k.hi; e.sig = k.sig; e.pad0 = 0; e.pad1 = 0; F.q[sl] = e; }
                else { F.log[sl] = ec_pk_tag(pc[0].codes, pc[0].len); n.perfect++; }
            } else generic = true;                                     // no room left: the generic queue takes the read
        }
        if (W.base != FX_NOBLOCK) W.used += cnt;
    }
    if (__any_sync(0xffffffffu, generic)) {
        if (generic) {
            GEntry ge; ge.seq_addr = (uint64_t)gseq; ge.seq_len = r;
            ge.qual_addr = (uint64_t)gqual; ge.qual_len = q;
            const uint32_t slot = atomicAdd(&F.St->g_count, 1u);
            if (slot < F.St->g_cap) F.gqueue[slot] = ge;
            else F.St->spec_fail = 1u;                                 // no room: the exact kernel redoes the chunk
        }
    }
}

// ---- Extract+Count: packed key table -------------------------------------------------------------------------------
// 16-byte slots {tag, count}: tag = ec_pk_tag(codes, len) (0 = empty), open addressing with linear probing; the host keeps
// the load factor <= 1/2 (f2q_api.cu: ec_reserve).  `add` occurrences of the key are counted.
__device__ __forceinline__ uint64_t ec_pk_hash(unsigned long long tag) {
    uint64_t h = tag * 0x9E3779B97F4A7C15ull;
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 29;
    return h;
}
__device__ __forceinline__ void ec_pk_insert(const EcTable& E, const Outputs& O, unsigned long long tag, unsigned long long add) {
    uint64_t i = ec_pk_hash(tag) & E.pk_mask;
    for (uint64_t probes = 0; probes <= E.pk_mask; probes++, i = (i + 1) & E.pk_mask) {
        unsigned long long cur = *(volatile unsigned long long*)(E.pk_slots + 2 * i);
        if (cur == 0ull) {
            cur = atomicCAS(E.pk_slots + 2 * i, 0ull, tag);
            if (cur == 0ull) { atomicAdd(E.pk_n, 1ull); cur = tag; }
        }
        if (cur == tag) { atomicAdd(E.pk_slots + 2 * i + 1, add); return; }
    }
    atomicOr(O.error, ERR_EC_FULL);
}

// the insert log of a chunk -> the packed table, when the chunk's speculation verified (DevState::spec_ok).  One thread per
// entry; equal tags inside a warp are added once (bar-seq abundances are heavy-tailed: the hot keys would serialise)
__device__ __forceinline__ uint32_t fx_queue_fill(const DevState* St, uint32_t q_cap) {
    const uint32_t full = q_cap - q_cap % FX_BLOCK;
    return min(St->q_count, full);
}
__global__ void __launch_bounds__(256) k_ec_commit(const DevState* St, EcTable E, Outputs O, const unsigned long long* __restrict__ log, uint32_t q_cap) {
    if (!St->spec_ok) return;
    const uint32_t n = fx_queue_fill(St, q_cap);
    constexpr int U = 4;                                               // entries per thread and round: U independent probes in flight
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride * U) {
        unsigned long long tag[U], cur[U];
        uint64_t slot[U];
        uint32_t add[U];
        #pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i = base + u * stride + threadIdx.x;
            tag[u] = i < n ? log[i] : 0ull;
        }
        #pragma unroll
        for (int u = 0; u < U; u++) {
            // equal tags inside a warp are added once
            const uint32_t peers = __match_any_sync(0xffffffffu, tag[u]);
            add[u] = (tag[u] && (uint32_t)(__ffs((int)peers) - 1) == (threadIdx.x & 31u)) ? (uint32_t)__popc(peers) : 0u;
            slot[u] = ec_pk_hash(tag[u]) & E.pk_mask;
            cur[u] = add[u] ? *(volatile unsigned long long*)(E.pk_slots + 2 * slot[u]) : 0ull;
        }
        #pragma unroll
        for (int u = 0; u < U; u++) {
            if (!add[u]) continue;
            if (cur[u] == tag[u]) atomicAdd(E.pk_slots + 2 * slot[u] + 1, (unsigned long long)add[u]);      // the usual case: the key is there, at its home slot
            else ec_pk_insert(E, O, tag[u], (unsigned long long)add[u]);
        }
    }
}

// re-insert every entry of an old packed table into a new (larger, zeroed) one
__global__ void __launch_bounds__(256) k_ec_rehash(const unsigned long long* __restrict__ old_slots, uint64_t old_cap, EcTable E, Outputs O) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < old_cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long tag = old_slots[2 * i];
        if (!tag) continue;
        uint64_t j = ec_pk_hash(tag) & E.pk_mask;
        for (;;) {
            const unsigned long long cur = atomicCAS(E.pk_slots + 2 * j, 0ull, tag);
            if (cur == 0ull) { E.pk_slots[2 * j + 1] = old_slots[2 * i + 1]; break; }
            j = (j + 1) & E.pk_mask;
        }
    }
}

// compacts the packed table into (tag, count) pairs for f2q_ec_drain / f2q_ec_merge; out_n counts them.  A block takes 1024
// slots per round and reserves the room for their keys with ONE atomic (an atomic with a return value per key — all on one
// address — ran at 20 M keys/s: 70 ms for the 1.4 M keys of the Bar-seq workload)
// (tags at tags[i * ts], counts at counts[i * cs]: the packed table interleaves them, the byte-arena table has two arrays)
__global__ void __launch_bounds__(256) k_ec_compact(const unsigned long long* __restrict__ tags, uint32_t ts, const unsigned long long* __restrict__ counts,
                                                    uint32_t cs, uint64_t cap, unsigned long long* __restrict__ out, unsigned long long* out_n) {
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint64_t base = (uint64_t)blockIdx.x * 1024u; base < cap; base += (uint64_t)gridDim.x * 1024u) {
        unsigned long long tag[4], cnt[4];
        uint32_t mine = 0;
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint64_t i = base + (uint64_t)k * 256u + threadIdx.x;
            tag[k] = i < cap ? tags[i * ts] : 0ull;
            cnt[k] = tag[k] ? counts[i * cs] : 0ull;
            mine += tag[k] ? 1u : 0u;
        }
        uint32_t incl = mine;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += y; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
        #pragma unroll
        for (int w = 0; w < 8; w++) { const uint32_t v = s_warp[w]; if ((uint32_t)w < warp) before += v; total += v; }
        if (threadIdx.x == 0 && total) s_base = atomicAdd(out_n, (unsigned long long)total);
        __syncthreads();
        if (mine) {
            unsigned long long o = s_base + before + incl - mine;
            #pragma unroll
            for (int k = 0; k < 4; k++) if (tag[k]) { out[2 * o] = tag[k]; out[2 * o + 1] = cnt[k]; o++; }
        }
        __syncthreads();
    }
}
#endif  // __CUDACC__

// ---- resolver of non-exact flex keys (fast2q.py:692-750 + 660-690 for keys of one shape) ---------------------------------
// Candidates of a key are the library keys of the same signature (the caller checked that the key's byte length has exactly
// this one shape in the library).  Pigeonhole: a library key within m mismatches agrees exactly with the read key on one of
// m+1 segments of its symbols, so the seed index (fxs_*) maps (signature, segment, segment value) -> the keys holding that
// value there.  G lanes work on one key: they share the bucket probes and split the candidates, then merge
// (min distance, how many attain it, which) with shuffles.  G is chosen by the host from the mean bucket size.
__host__ __device__ __forceinline__ uint64_t fxs_tag(uint32_t sig, uint32_t seg, uint64_t value) {
    return (1ull << 63) | ((uint64_t)(sig & 0x3FFFu) << 48) | ((uint64_t)(seg & 15u) << 44) | value;     // value < 2^40 (a segment is <= 20 symbols)
}
__host__ __device__ __forceinline__ uint32_t fxs_hash(uint64_t tag) {
    uint64_t h = tag * 0x9E3779B97F4A7C15ull;
    h ^= h >> 31; h *= 0xD6E8FEB86659FD93ull;
    return (uint32_t)(h >> 32);
}
// 2-bit codes of symbols [b0, b1) of an 80-bit key (b1 - b0 <= 20)
__host__ __device__ __forceinline__ uint64_t fx_segment(uint64_t lo, uint32_t hi, uint32_t b0, uint32_t b1) {
    const uint32_t sh = 2u * b0, nb = 2u * (b1 - b0);
    uint64_t v = sh >= 64u ? (uint64_t)hi >> (sh - 64u) : (lo >> sh) | (sh ? (uint64_t)hi << (64u - sh) : 0ull);
    return nb >= 64u ? v : v & ((1ull << nb) - 1ull);
}
// even bits of symbols [b0, b1) of the 80-bit difference mask (lo part, hi part)
__host__ __device__ __forceinline__ void fx_even_range(uint32_t b0, uint32_t b1, uint64_t& mlo, uint32_t& mhi) {
    auto below = [](uint32_t b, uint64_t& lo, uint32_t& hi) {          // even bits of symbols [0, b)
        lo = b >= 32u ? 0x5555555555555555ull : (((1ull << (2u * b)) - 1ull) & 0x5555555555555555ull);
        hi = b <= 32u ? 0u : (((1u << (2u * (b - 32u))) - 1u) & 0x55555555u);
    };
    uint64_t al, bl; uint32_t ah, bh;
    below(b0, al, ah); below(b1, bl, bh);
    mlo = bl & ~al; mhi = bh & ~ah;
}

#ifdef __CUDACC__
struct FxBest { int d; uint32_t n, idx; };

template <int G>
__device__ __forceinline__ uint32_t fx_resolve_group(const LibTables& T, int m, const FlexQ& e, uint32_t gl /*lane in group*/) {
    const uint32_t nsym = fx_sig_symbols(e.sig), parts = T.fxs_parts;
    const int nbad = __popcll(e.bad);
    FxBest b{m + 1, 0u, 0u};
    if (nbad <= m && nsym >= parts) {
        // difference bits forced by the bad symbols
        const uint64_t bad_lo = spread_even((uint32_t)e.bad);
        const uint32_t bad_hi = (uint32_t)spread_even((uint32_t)(e.bad >> 32));
        for (uint32_t s = 0; s < parts; s++) {
            const uint32_t b0 = s * nsym / parts, b1 = (s + 1) * nsym / parts;
            uint64_t slo; uint32_t shi;
            fx_even_range(b0, b1, slo, shi);
            if ((bad_lo & slo) | (uint64_t)(bad_hi & shi)) continue;   // a symbol that is not A/C/G/T never agrees exactly
            const uint64_t tag = fxs_tag(e.sig, s, fx_segment(e.lo, e.hi, b0, b1));
            uint32_t h = fxs_hash(tag) & T.fxs_mask, start = 0, count = 0;
            for (;;) {
                const uint4 raw = __ldg(T.fxs_slots + h);
                const uint64_t t = ((uint64_t)raw.y << 32) | raw.x;
                if (t == 0) break;
                if (t == tag) { start = raw.z; count = raw.w; break; }
                h = (h + 1) & T.fxs_mask;
            }
            for (uint32_t c = gl; c < count; c += (uint32_t)G) {
                const uint4 it = __ldg(T.fxs_recs + start + c);        // {lo.x, lo.y, hi | feature index << 16 (hi is 16 bits), -}
                const uint64_t xl = e.lo ^ (((uint64_t)it.y << 32) | it.x);
                const uint32_t xh = e.hi ^ (it.z & 0xFFFFu);
                const uint64_t dl = ((xl | (xl >> 1)) & 0x5555555555555555ull) | bad_lo;
                const uint32_t dh = ((xh | (xh >> 1)) & 0x5555u) | bad_hi;
                bool dup = false;                                      // already met through an earlier agreeing segment?
                for (uint32_t s2 = 0; s2 < s; s2++) {
                    uint64_t ml; uint32_t mh;
                    fx_even_range(s2 * nsym / parts, (s2 + 1) * nsym / parts, ml, mh);
                    if (((dl & ml) | (uint64_t)(dh & mh)) == 0) { dup = true; break; }
                }
                if (dup) continue;
                const int dist = __popcll(dl) + __popc(dh);
                if (dist < b.d) { b.d = dist; b.n = 1; b.idx = it.w; }
                else if (dist == b.d && dist <= m) b.n++;
            }
        }
    }
    // merge over the group
    #pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const int od = __shfl_xor_sync(0xffffffffu, b.d, o);
        const uint32_t on = __shfl_xor_sync(0xffffffffu, b.n, o), oi = __shfl_xor_sync(0xffffffffu, b.idx, o);
        if (od < b.d) { b.d = od; b.n = on; b.idx = oi; }
        else if (od == b.d) b.n += on;
    }
    return (b.d <= m && b.n == 1) ? b.idx : RES_NONE;
}

template <int G>
__global__ void __launch_bounds__(256) k_resolve_flex(LibTables T, int m, const DevState* St, const FlexQ* __restrict__ q, uint32_t q_cap,
                                                      unsigned long long* counts, unsigned long long* stats, unsigned long long* memo_stats) {
    const uint32_t gl = threadIdx.x % G, grp = threadIdx.x / G, groups = blockDim.x / G;
    uint32_t imperfect = 0, nonal = 0;
    const uint32_t n = St->spec_ok ? fx_queue_fill(St, q_cap) : 0u;     // (a chunk the exact kernel re-parsed left nothing valid here)
    // (every thread of a warp runs the same number of rounds: the shuffles inside need the whole warp)
    const uint32_t per_round = gridDim.x * groups;
    for (uint32_t base = 0; base < n; base += per_round) {
        const uint32_t i = base + blockIdx.x * groups + grp;
        FlexQ e;
        e.lo = 0; e.bad = 0; e.hi = 0; e.sig = 0; e.pad0 = 0; e.pad1 = 0;
        if (i < n) e = q[i];
        const bool act = e.sig != 0;                                    // (sig 0: the unused rest of a warp's block)
        // memo first (keys without bad symbols): entry = {lo.x, lo.y, hi | sig << 16, result + 2}
        const uint32_t mz = e.hi | (e.sig << 16);
        const bool memo_ok = act && T.memo && e.bad == 0;
        uint32_t cached = 0;
        if (memo_ok && gl == 0) cached = memo_lookup(T, (uint32_t)e.lo, (uint32_t)(e.lo >> 32), mz);
        if (G > 1) cached = __shfl_sync(0xffffffffu, cached, (threadIdx.x & 31u) - gl);
        if (memo_ok && gl == 0) { atomicAdd(memo_stats, 1ull); if (cached) atomicAdd(memo_stats + 1, 1ull); }
        if (!act || cached) e.bad = ~0ull;                              // too many bad symbols: no work
        uint32_t r = fx_resolve_group<G>(T, m, e, gl);
        if (cached) r = cached - 2u;
        else if (memo_ok && gl == 0) memo_store(T, (uint32_t)e.lo, (uint32_t)(e.lo >> 32), mz, r + 2u);
        if (act && gl == 0) { if (r != RES_NONE) { atomicAdd(counts + r, 1ull); imperfect++; } else nonal++; }
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
