// tile.cuh — the fused hot kernel: ONE pass over the FASTQ bytes.
//
//   K1 line scan      newline flags per 16-byte chunk, block scan, decoupled look-back across tiles for the
//                     line phase (records are "every 4 lines from byte 0", fast2q.py:324-328)
//   K2 extract        fixed-position window with Python slice clamping, rstrip, Phred fail-set test (fast2q.py:349-360)
//   K4 lookup/count   2-bit pack, exact probe of the packed-key table, shared-memory histogram (fast2q.py:365-367)
//   -> non-exact keys go to the resolver queue (K5), undecidable reads to the generic queue.
//
// Work decomposition: persistent CTAs take tiles by atomic ticket (so a tile's predecessors are always running or
// done — the look-back cannot deadlock).  A tile is TILE_ROWS x 128 B in shared memory with the 128-byte XOR
// swizzle; thread t scans row t.  A read belongs to the tile that holds the newline ending its header line.
#pragma once

#include "f2q_dev.cuh"
#include "generic.cuh"
#include "resolve.cuh"

namespace f2q {

enum { POLICY_GENERIC = 0, POLICY_FAST1 = 1 };

struct TileParams {
    const uint8_t* buf;        // 128-byte aligned base of the chunk buffer
    DevState* S;
    uint32_t* status;          // look-back status words, zeroed before the launch
    uint32_t* ticket;          // tile ticket counter, zeroed before the launch
    int stitch;                // 1: parse [0, S->stitch_len) of the carry buffer; 0: parse [S->beg, S->end)
    QEntry* queue;
    GEntry* gqueue;
    uint32_t hist_smem;        // 1: per-CTA shared-memory histogram of n_keys u32
};

constexpr int QSTAGE = TILE_THREADS;     // staged queue entries per read-loop iteration
constexpr uint32_t LB_SPIN_LIMIT = 1u << 24;

__host__ __device__ inline size_t tile_smem_bytes(uint32_t hist_entries) {
    return (size_t)TILE_BYTES + 128 + NL_CAP * 2 + QSTAGE * sizeof(QEntry) + (size_t)hist_entries * 4;
}

__device__ __forceinline__ uint32_t tile_byte(const uint8_t* tile, uint32_t o) { return tile[swz(o)]; }
__device__ __forceinline__ uint32_t tile_word(const uint8_t* tile, uint32_t o /*4-aligned*/) {
    return *reinterpret_cast<const uint32_t*>(tile + swz(o));
}

// 4 bytes at an arbitrary tile offset, little endian
struct WordReader {
    const uint8_t* tile; uint32_t a, sh, prev;
    __device__ __forceinline__ WordReader(const uint8_t* t, uint32_t o) : tile(t), a(o & ~3u), sh((o & 3u) * 8u) { prev = tile_word(tile, a); }
    __device__ __forceinline__ uint32_t next() {
        a += 4;
        uint32_t nx = tile_word(tile, a);
        uint32_t w = __funnelshift_r(prev, nx, sh);
        prev = nx;
        return w;
    }
};

// any byte b of tile[o, o+n) with 33 <= b <= fmax ?   (fast2q.py:357 with the fail set of :1127)
__device__ __forceinline__ bool qual_fails_tile(const uint8_t* tile, uint32_t o, int n, int fmax) {
    if (fmax == 0 || n <= 0) return false;
    const uint32_t add_ge = (0x80u - 33u) * 0x01010101u, add_gt = (0x80u - (uint32_t)(fmax + 1)) * 0x01010101u;
    WordReader rd(tile, o);
    uint32_t acc = 0;
    for (int k = 0; k < n; k += 4) {
        uint32_t w = rd.next();
        if (n - k < 4) w &= (1u << (8 * (n - k))) - 1u;               // bytes past the slice become 0 (never fail)
        uint32_t lo7 = w & 0x7F7F7F7Fu;
        uint32_t ge33 = ((lo7 + add_ge) | w);                          // bit 7: byte >= 33
        uint32_t gtmax = ((lo7 + add_gt) | w);                         // bit 7: byte >  fmax
        acc |= ge33 & ~gtmax;
    }
    return (acc & 0x80808080u) != 0;
}

// 2-bit pack of tile[o, o+n), n <= 32.  bad = mask of symbols outside ACGT (after upper()); their key bits are 0
__device__ __forceinline__ void pack_tile(const uint8_t* tile, uint32_t o, int n, uint64_t& key, uint32_t& bad) {
    key = 0; bad = 0;
    if (n <= 0) return;
    WordReader rd(tile, o);
    uint32_t anybad = 0;
    for (int k = 0; k < n; k += 4) {
        uint32_t w = rd.next();
        if (n - k < 4) { uint32_t keep = (1u << (8 * (n - k))) - 1u; w = (w & keep) | (0x41414141u & ~keep); }   // pad with 'A' (code 0)
        uint32_t codes = (w >> 1) & 0x03030303u;
        uint32_t t = (codes | (codes >> 4)) & 0x00330033u;
        t = (t | (t >> 8)) & 0x3333u;                                  // nibble i = code of byte i
        uint32_t expect = __byte_perm(0x47544341u, 0u, t);              // code -> 'A','C','T','G'
        anybad |= ~eq_bytes(w & 0xDFDFDFDFu, expect) & 0x80808080u;
        uint32_t p = (t | (t >> 2)) & 0x0F0Fu;
        p = (p | (p >> 4)) & 0xFFu;                                    // 4 symbols -> 8 bits
        key |= (uint64_t)p << (2 * k);
    }
    if (anybad) {                                                       // rare: redo byte-wise for the exact mask
        key = 0;
        for (int k = 0; k < n; k++) {
            uint32_t code, c = tile_byte(tile, o + k);
            if (base_code(c, code)) key |= (uint64_t)code << (2 * k); else bad |= 1u << k;
        }
    }
}

__device__ __noinline__ uint64_t find_newline_global(const uint8_t* buf, uint64_t from, uint64_t end) {
    for (uint64_t p = from; p < end; p++) if (buf[p] == '\n') return p;
    return end;
}

// per-thread accumulators, reduced once per CTA
struct Acc {
    unsigned long long reads, perfect, imperfect, nonal, qfail, last_end;
};

template <int POLICY>
__global__ void __launch_bounds__(TILE_THREADS) k_tile(TileParams P, const GenericCfg* __restrict__ Gp, LibTables T, EcTable E, Outputs O) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tile = smem;
    uint16_t* nlpos = reinterpret_cast<uint16_t*>(smem + TILE_BYTES + 128);
    QEntry* qstage = reinterpret_cast<QEntry*>(smem + TILE_BYTES + 128 + NL_CAP * 2);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + TILE_BYTES + 128 + NL_CAP * 2 + QSTAGE * sizeof(QEntry));
    __shared__ uint32_t s_wsum[TILE_THREADS / 32];
    __shared__ uint32_t s_tile, s_p0, s_A, s_qn, s_qbase;
    __shared__ GenericCfg s_G;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevState* S = P.S;
    const uint64_t beg = P.stitch ? 0ull : S->beg;
    const uint64_t end = P.stitch ? (uint64_t)S->stitch_len : S->end;
    const bool eof = P.stitch ? (S->stitch_eof != 0) : (S->is_last != 0);
    if (end <= beg) return;

    // configuration into shared memory (the generic path reads it through a pointer)
    for (uint32_t i = tid; i < sizeof(GenericCfg) / 4; i += TILE_THREADS)
        reinterpret_cast<uint32_t*>(&s_G)[i] = reinterpret_cast<const uint32_t*>(Gp)[i];
    if (P.hist_smem) for (uint32_t i = tid; i < T.n_keys; i += TILE_THREADS) hist[i] = 0;
    __syncthreads();
    const DevCfg& C = s_G.c;

    const uint64_t first_tile = beg / OWN_BYTES;
    const uint64_t n_tiles = (end - 1) / OWN_BYTES + 1;
    const uint8_t* __restrict__ buf = P.buf;
    Acc acc{0, 0, 0, 0, 0, 0};
    unsigned long long gst[F2Q_N_STATS] = {0, 0, 0, 0, 0};            // stats of reads handled inline by the generic code

    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(P.ticket, 1u);
        __syncthreads();
        const uint64_t t = first_tile + s_tile;
        if (t >= n_tiles) break;
        const uint64_t base = t * OWN_BYTES;

        // ---- load the tile (coalesced 16-byte loads -> swizzled shared memory); bytes outside [beg,end) become 0 ----
        #pragma unroll
        for (int j = 0; j < TILE_BYTES / 16 / TILE_THREADS; j++) {
            const uint32_t c = j * TILE_THREADS + tid;
            const uint64_t g = base + (uint64_t)c * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (g + 16 > beg && g < end) {
                v = ldg_stream(reinterpret_cast<const uint4*>(buf + g));
                if (g < beg || g + 16 > end) {                          // partial chunk at either end of the range
                    uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    for (int b = 0; b < 16; b++) {
                        uint64_t pos = g + b;
                        if (pos < beg || pos >= end) w[b >> 2] &= ~(0xFFu << (8 * (b & 3)));
                    }
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            *reinterpret_cast<uint4*>(tile + swz(c * 16)) = v;
        }
        __syncthreads();

        // ---- K1: newline flags of row tid.  comb[j] bit (8*i + k) <-> byte (4*k + i) of chunk j ----
        uint32_t comb[8];
        uint32_t cnt = 0;
        #pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint4 v = *reinterpret_cast<const uint4*>(tile + tid * ROW_BYTES + ((j ^ (tid & 7)) << 4));
            const uint32_t z0 = eq_bytes(v.x, 0x0A0A0A0Au), z1 = eq_bytes(v.y, 0x0A0A0A0Au);
            const uint32_t z2 = eq_bytes(v.z, 0x0A0A0A0Au), z3 = eq_bytes(v.w, 0x0A0A0A0Au);
            comb[j] = (z0 >> 7) | (z1 >> 6) | (z2 >> 5) | (z3 >> 4);
            cnt += __popc(comb[j]);
        }
        // block exclusive scan of cnt (rank of this row's first newline within the tile)
        uint32_t incl = cnt;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
        #pragma unroll
        for (int w = 0; w < TILE_THREADS / 32; w++) { uint32_t x = s_wsum[w]; if (w < (int)warp) wbase += x; total += x; }
        const uint32_t excl = wbase + incl - cnt;

        // ---- decoupled look-back (warp that holds row OWN_ROWS knows the owned newline count A) ----
        if (warp == OWN_ROWS / 32) {
            const uint32_t A = __shfl_sync(0xffffffffu, excl, OWN_ROWS % 32);
            const uint64_t rel = t - first_tile;
            uint32_t p0 = 0;
            if (rel != 0) {
                if (lane == 0) st_volatile_u32(P.status + rel, LB_FLAG_AGG | (A & LB_VALUE_MASK));
                int64_t look = (int64_t)rel - 1;
                for (;;) {
                    const int64_t idx = look - lane;
                    uint32_t sw = LB_FLAG_PREFIX;                       // before the first tile: prefix 0
                    if (idx >= 0) {
                        uint32_t spins = 0;
                        do { sw = ld_volatile_u32(P.status + idx); } while ((sw >> 30) == 0 && ++spins < LB_SPIN_LIMIT);
                        if ((sw >> 30) == 0) { atomicOr(O.error, ERR_LOOKBACK_TIMEOUT); sw = LB_FLAG_PREFIX; }
                    }
                    const uint32_t is_prefix = __ballot_sync(0xffffffffu, (sw >> 30) == 2);
                    const int first = is_prefix ? __ffs(is_prefix) - 1 : 32;
                    const uint32_t contrib = ((int)lane <= first) ? (sw & LB_VALUE_MASK) : 0u;
                    p0 += __reduce_add_sync(0xffffffffu, contrib);
                    if (is_prefix) break;
                    look -= 32;
                }
            }
            if (lane == 0) {
                st_volatile_u32(P.status + rel, LB_FLAG_PREFIX | ((p0 + A) & LB_VALUE_MASK));
                s_p0 = p0;
                if (t == n_tiles - 1 && !P.stitch) S->nl_total = (p0 + A) & LB_VALUE_MASK;
            }
        }

        // A (owned newline count) for everybody: ranks < A lie in owned rows
        if (tid == OWN_ROWS) s_A = excl;

        // ---- passes over windows of NL_CAP newline ranks (one pass unless lines are shorter than ~8 bytes) ----
        for (uint32_t pass_base = 0;; pass_base += NL_CAP - 3) {
            // emit the positions of this row's newlines
            {
                uint32_t r = excl;
                #pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint32_t m = comb[j];
                    if (m == 0) continue;
                    if ((m & (m - 1)) == 0) {                           // single newline in the chunk (the common case)
                        const uint32_t b = __ffs(m) - 1;
                        const uint32_t pos = tid * ROW_BYTES + j * 16 + 4 * (b & 7) + (b >> 3);
                        const uint32_t rr = r - pass_base;
                        if (rr < NL_CAP) nlpos[rr] = (uint16_t)pos;
                        r++;
                    } else {
                        for (int byte = 0; byte < 16; byte++) {
                            if ((m >> (8 * (byte & 3) + (byte >> 2))) & 1u) {
                                const uint32_t rr = r - pass_base;
                                if (rr < NL_CAP) nlpos[rr] = (uint16_t)(tid * ROW_BYTES + j * 16 + byte);
                                r++;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            const uint32_t p0 = s_p0;
            const uint32_t A = s_A;
            // ranks q with (p0 + q) % 4 == 0 end a header line; read j of the tile has q = q0 + 4j
            const uint32_t q0 = (4u - (p0 & 3u)) & 3u;
            const uint32_t R = (A > q0) ? (A - q0 + 3u) / 4u : 0u;
            // reads of this pass: pass_base <= q and q + 3 < pass_base + NL_CAP
            uint32_t j_lo = 0;
            if (pass_base > q0) j_lo = (pass_base - q0 + 3u) / 4u;
            uint32_t j_hi = R;                                          // exclusive
            const bool last_pass = (pass_base + NL_CAP >= total);
            if (!last_pass) {
                const uint32_t lim = pass_base + NL_CAP - 3;            // q must be < lim
                if (lim > q0) { uint32_t jh = (lim - q0 + 3u) / 4u; if (jh < j_hi) j_hi = jh; } else j_hi = 0;
            }
            const uint32_t region_end = (uint32_t)min((uint64_t)TILE_BYTES, end - base);   // valid bytes of the tile region
            const bool region_has_eof = (base + TILE_BYTES >= end);

            for (uint32_t jb = j_lo; jb < j_hi; jb += TILE_THREADS) {
                if (tid == 0) s_qn = 0;
                __syncthreads();
                const uint32_t j = jb + tid;
                if (j < j_hi) {
                    const uint32_t q = q0 + 4u * j;                     // rank of the newline that ends the header line
                    const uint32_t qi = q - pass_base;
                    // line geometry (tile offsets); spill = some line end is outside shared memory
                    uint32_t s0 = (uint32_t)nlpos[qi] + 1u, e0 = 0, s3 = 0, e3 = 0;
                    bool complete = false, spill = false;
                    if (q + 3 < total) {
                        e0 = nlpos[qi + 1]; s3 = (uint32_t)nlpos[qi + 2] + 1u; e3 = nlpos[qi + 3];
                        complete = true;
                        acc.last_end = max(acc.last_end, (unsigned long long)(base + e3 + 1));
                    } else if (region_has_eof) {
                        // no further newline exists: only an unterminated final quality line can complete the record
                        if (eof && q + 2 < total) {
                            e0 = nlpos[qi + 1]; s3 = (uint32_t)nlpos[qi + 2] + 1u; e3 = region_end;
                            if (s3 < region_end) { complete = true; acc.last_end = max(acc.last_end, (unsigned long long)end); }
                        }
                    } else spill = true;

                    if (spill) {
                        // finish the geometry in global memory, then hand the read to the generic code
                        uint64_t pos[4]; uint32_t have = total - q;      // newlines q .. q+have-1 are in shared memory (1..3)
                        for (uint32_t k = 0; k < 4; k++) pos[k] = (k < have) ? base + nlpos[qi + k] : 0;
                        uint64_t from = base + TILE_BYTES;
                        bool ok = true;
                        for (uint32_t k = have; k < 4; k++) {
                            uint64_t p = find_newline_global(buf, from, end);
                            if (p >= end) {
                                if (k == 3 && eof && from < end) { pos[3] = end; }     // unterminated final line
                                else ok = false;
                                break;
                            }
                            pos[k] = p; from = p + 1;
                        }
                        if (ok) {
                            acc.reads++;
                            acc.last_end = max(acc.last_end, (unsigned long long)min(pos[3] + 1, end));
                            GEntry ge; ge.seq_addr = (uint64_t)(buf + pos[0] + 1); ge.seq_len = (uint32_t)(pos[1] - pos[0] - 1);
                            ge.qual_addr = (uint64_t)(buf + pos[2] + 1); ge.qual_len = (uint32_t)(pos[3] - pos[2] - 1);
                            uint32_t slot = (POLICY == POLICY_GENERIC) ? 0xFFFFFFFFu : atomicAdd(&S->g_count, 1u);
                            if (slot < S->g_cap) P.gqueue[slot] = ge;
                            else {
                                const uint8_t* Rp = (const uint8_t*)ge.seq_addr; const uint8_t* Qp = (const uint8_t*)ge.qual_addr;
                                g_process_read(s_G, T, E, O, Rp, g_rstrip(Rp, (int)ge.seq_len), Qp, g_rstrip(Qp, (int)ge.qual_len), gst);
                            }
                        }
                    } else if (complete) {
                        acc.reads++;
                        if (POLICY == POLICY_GENERIC) {
                            const uint8_t* Rp = buf + base + s0; const uint8_t* Qp = buf + base + s3;
                            g_process_read(s_G, T, E, O, Rp, g_rstrip(Rp, (int)(e0 - s0)), Qp, g_rstrip(Qp, (int)(e3 - s3)), gst);
                        } else {
                            // ---- K2: rstrip, window, Phred test ----
                            while (e0 > s0 && is_py_space(tile_byte(tile, e0 - 1))) e0--;
                            while (e3 > s3 && is_py_space(tile_byte(tile, e3 - 1))) e3--;
                            int lo, hi, qlo, qhi;
                            py_slice((int)(e0 - s0), C.starts[0], C.starts[0] + C.length, lo, hi);
                            py_slice((int)(e3 - s3), C.starts[0], C.starts[0] + C.length, qlo, qhi);
                            if (qual_fails_tile(tile, s3 + qlo, qhi - qlo, C.fmax_ph)) acc.qfail++;
                            else {
                                // ---- K4: pack, exact lookup, count ----
                                uint64_t key; uint32_t bad; const uint32_t klen = (uint32_t)(hi - lo);
                                pack_tile(tile, s0 + lo, (int)klen, key, bad);
                                const bool generic_len = (T.generic_len_mask >> min(klen, 63u)) & 1ull;
                                uint32_t idx = SLOT_EMPTY;
                                if (!generic_len && bad == 0) idx = fast_lookup(T, key, klen);
                                if (idx != SLOT_EMPTY) {
                                    acc.perfect++;
                                    if (P.hist_smem) atomicAdd(hist + idx, 1u);
                                    else atomicAdd(O.counts + idx, 1ull);
                                } else if (generic_len) {
                                    // library keys of this length exist that the packed tables cannot hold
                                    GEntry ge; ge.seq_addr = (uint64_t)(buf + base + s0); ge.seq_len = e0 - s0;
                                    ge.qual_addr = (uint64_t)(buf + base + s3); ge.qual_len = e3 - s3;
                                    uint32_t slot = atomicAdd(&S->g_count, 1u);
                                    if (slot < S->g_cap) P.gqueue[slot] = ge;
                                    else {
                                        g_process_read(s_G, T, E, O, (const uint8_t*)ge.seq_addr, (int)ge.seq_len, (const uint8_t*)ge.qual_addr, (int)ge.qual_len, gst);
                                    }
                                } else if (C.miss <= 0) acc.nonal++;
                                else {
                                    const uint32_t sl = atomicAdd(&s_qn, 1u);
                                    qstage[sl].key = key; qstage[sl].bad = bad; qstage[sl].len = klen;
                                }
                            }
                        }
                    }
                }
                // ---- flush the staged non-exact keys: one global atomic per iteration ----
                __syncthreads();
                const uint32_t nq = s_qn;
                if (nq) {
                    if (tid == 0) s_qbase = atomicAdd(&S->q_count, nq);
                    __syncthreads();
                    const uint32_t qb = s_qbase;
                    if (tid < nq) {
                        if (qb + tid < S->q_cap) P.queue[qb + tid] = qstage[tid];
                        else {                                           // queue full: resolve right here
                            const QEntry e = qstage[tid];
                            const uint32_t r = resolve_thread(T, C.miss, e.key, e.bad, e.len);
                            if (r != RES_NONE) { atomicAdd(O.counts + r, 1ull); acc.imperfect++; } else acc.nonal++;
                        }
                    }
                }
            }
            if (last_pass) break;
            __syncthreads();
        }
    }

    // ---- CTA epilogue: histogram and statistics ----
    __syncthreads();
    if (P.hist_smem)
        for (uint32_t i = tid; i < T.n_keys; i += TILE_THREADS) { uint32_t v = hist[i]; if (v) atomicAdd(O.counts + i, (unsigned long long)v); }
    acc.perfect += gst[F2Q_STAT_PERFECT]; acc.imperfect += gst[F2Q_STAT_IMPERFECT];
    acc.nonal += gst[F2Q_STAT_NON_ALIGNED]; acc.qfail += gst[F2Q_STAT_QUALITY_FAILED];
    unsigned long long v[5] = {acc.reads, acc.perfect, acc.imperfect, acc.nonal, acc.qfail};
    #pragma unroll
    for (int k = 0; k < 5; k++) {
        unsigned long long x = v[k];
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xffffffffu, x, d);
        if (lane == 0 && x) atomicAdd(O.stats + k, x);
    }
    unsigned long long le = acc.last_end;
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) le = max(le, __shfl_down_sync(0xffffffffu, le, d));
    if (lane == 0 && le && !P.stitch) atomicMax(&S->last_rec_end, le);
}

}  // namespace f2q
