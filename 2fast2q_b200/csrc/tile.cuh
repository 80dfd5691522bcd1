// tile.cuh — the fused hot kernel: ONE pass over the FASTQ bytes.
//
//   K1 line scan      per-row newline bitmasks (SWAR compare + dp4a gather), block scan of the counts, decoupled
//                     look-back across tiles for the line phase (records are "every 4 lines from byte 0",
//                     fast2q.py:324-328); rows exchange their masks through shared memory
//   K2 extract        fixed-position window with Python slice clamping, rstrip, Phred fail-set test (fast2q.py:349-360)
//   K4 lookup/count   2-bit pack, exact probe of the packed-key table, shared-memory histogram (fast2q.py:365-367)
//   -> non-exact keys go to this CTA's segment of the resolver queue (K5), undecidable reads to the generic queue.
//
// Work decomposition: persistent CTAs take tiles by atomic ticket WHEN THEY ARE READY for them and publish the tile's
// newline count as soon as it is scanned (a tile's predecessors are always running or done, so the look-back cannot
// deadlock and rarely waits); latency is hidden by several small CTAs per SM rather than by pipelining inside a CTA.
// A tile is (NT + halo) rows of S = 16*CH bytes in shared memory, NT = threads per CTA, CH odd so that
// "thread t scans row t with 16-byte loads" is bank-conflict free without a swizzle; CH is chosen by the host so that
// S is just below the record length (about one read start per row).  Interior tiles are fetched by one TMA bulk copy
// (cp.async.bulk + mbarrier); the first/last tile of a range, which needs byte masking, is loaded by the threads.
// A read belongs to the row (thread) that holds the newline ending its header line.
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
    QEntry* queue;             // grid segments of seg_cap entries each
    uint32_t* seg_count;       // entries written per segment (one per CTA)
    uint32_t seg_cap;          // 0: resolve every non-exact key in place
    GEntry* gqueue;
    uint32_t hist_smem;        // 1: per-CTA shared-memory histogram of n_keys u32
    uint32_t halo_rows;        // read-ahead rows loaded and scanned behind the NT owned rows (2 .. TileGeom::HALO)
};

constexpr uint32_t LB_SPIN_LIMIT = 1u << 24;

template <int CH, int NT_>
struct TileGeom {
    static constexpr int S = CH * 16;                          // bytes per row
    static constexpr int NT = NT_;                             // owned rows = threads per CTA
    static constexpr int HALO = (1024 + S - 1) / S + 1;        // max read-ahead rows (>= 1 KiB); P.halo_rows of them are used
    static constexpr int ROWS = NT + HALO;
    static constexpr int OWN_BYTES = NT * S;
    static constexpr int BUF_BYTES = ROWS * S + 16;            // tile buffer (+16 so that word reads may run past the end)
    static constexpr int MASK_OFF = BUF_BYTES;                 // ROWS uint4 row masks
    static constexpr int HIST_OFF = MASK_OFF + ROWS * 16;
};

template <int CH, int NT>
__host__ __device__ inline size_t tile_smem_bytes(uint32_t hist_entries) {
    return (size_t)TileGeom<CH, NT>::HIST_OFF + (size_t)hist_entries * 4;
}

// ---- mbarrier / TMA bulk copy (sm_90+ PTX) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy by the TMA unit; completion is signalled on the mbarrier (bytes % 16 == 0, 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // earlier generic-proxy accesses to dst are ordered first
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- byte access to the (unswizzled) tile ---------------------------------------------------------------
struct WordReader {                 // 4 bytes at an arbitrary tile offset, little endian
    const uint8_t* tile; uint32_t a, sh, prev;
    __device__ __forceinline__ WordReader(const uint8_t* t, uint32_t o) : tile(t), a(o & ~3u), sh((o & 3u) * 8u) {
        prev = *reinterpret_cast<const uint32_t*>(tile + a);
    }
    __device__ __forceinline__ uint32_t next() {
        a += 4;
        uint32_t nx = *reinterpret_cast<const uint32_t*>(tile + a);
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
    auto test = [&](uint32_t w) {
        const uint32_t lo7 = w & 0x7F7F7F7Fu;
        const uint32_t ge33 = ((lo7 + add_ge) | w);                    // bit 7: byte >= 33
        const uint32_t gtmax = ((lo7 + add_gt) | w);                   // bit 7: byte >  fmax
        acc |= ge33 & ~gtmax;
    };
    const int full = n >> 2;
    for (int k = 0; k < full; k++) test(rd.next());
    if (n & 3) test(rd.next() & ((1u << (8 * (n & 3))) - 1u));         // bytes past the slice become 0 (never fail)
    return (acc & 0x80808080u) != 0;
}

// 2-bit pack of tile[o, o+n), n <= 32.  bad = mask of symbols outside ACGT (after upper()); their key bits are 0
__device__ __forceinline__ void pack_tile(const uint8_t* tile, uint32_t o, int n, uint64_t& key, uint32_t& bad) {
    key = 0; bad = 0;
    if (n <= 0) return;
    WordReader rd(tile, o);
    const int words = (n + 3) >> 2;
    for (int kw = 0; kw < words; kw++) {
        const int k = kw * 4;
        uint32_t w = rd.next();
        if (kw == words - 1 && (n & 3)) { const uint32_t keep = (1u << (8 * (n & 3))) - 1u; w = (w & keep) | (0x41414141u & ~keep); }   // pad with 'A' (code 0)
        uint32_t codes = (w >> 1) & 0x03030303u;
        uint32_t t = (codes | (codes >> 4)) & 0x00330033u;
        t = (t | (t >> 8)) & 0x3333u;                                  // nibble i = code of byte i
        const uint32_t expect = __byte_perm(0x47544341u, 0u, t);        // code -> 'A','C','T','G'
        const uint32_t ok = eq_bytes(w & 0xDFDFDFDFu, expect);          // 0x80 per valid byte
        if (ok != 0x80808080u) {                                        // some symbol is not A/C/G/T: flag it, zero its bits
            const uint32_t nb = (ok ^ 0x80808080u);
            bad |= ((nb * 0x00204081u) >> 28) << k;
            const uint32_t keep2 = (ok >> 7) * 3u;                      // 0x03 per valid byte
            codes &= keep2;
            t = (codes | (codes >> 4)) & 0x00330033u;
            t = (t | (t >> 8)) & 0x3333u;
        }
        uint32_t p = (t | (t >> 2)) & 0x0F0Fu;
        p = (p | (p >> 4)) & 0xFFu;                                    // 4 symbols -> 8 bits
        key |= (uint64_t)p << (2 * k);
    }
}

__device__ __noinline__ uint64_t find_newline_global(const uint8_t* buf, uint64_t from, uint64_t end) {
    for (uint64_t p = from; p < end; p++) if (buf[p] == '\n') return p;
    return end;
}

// 128-bit row mask helpers (bit b = byte b of the row is '\n')
struct Mask128 {
    uint64_t lo, hi;
    __device__ __forceinline__ bool empty() const { return (lo | hi) == 0; }
    __device__ __forceinline__ int pop_lowest() {               // index of the lowest set bit, which is cleared
        if (lo) { int b = __ffsll((long long)lo) - 1; lo &= lo - 1; return b; }
        int b = 64 + __ffsll((long long)hi) - 1; hi &= hi - 1; return b;
    }
};

// per-thread accumulators, reduced once per CTA
struct Acc {
    unsigned long long reads, perfect, imperfect, nonal, qfail, last_end;
};

template <int POLICY, int CH, int NT>
__global__ void __launch_bounds__(NT, POLICY == POLICY_FAST1 ? (768 / NT) : (512 / NT))
k_tile(TileParams P, const GenericCfg* __restrict__ Gp, LibTables T, EcTable E, Outputs O) {
    using G_ = TileGeom<CH, NT>;
    constexpr int S = G_::S;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* tile = smem;
    uint4* rowmask = reinterpret_cast<uint4*>(smem + G_::MASK_OFF);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + G_::HIST_OFF);
    __shared__ uint32_t s_wsum[NT / 32];
    __shared__ uint32_t s_tile, s_p0, s_qn;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ GenericCfg s_G;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevState* St = P.S;
    const uint64_t beg = P.stitch ? 0ull : St->beg;
    const uint64_t end = P.stitch ? (uint64_t)St->stitch_len : St->end;
    const bool eof = P.stitch ? (St->stitch_eof != 0) : (St->is_last != 0);
    if (end <= beg) { if (tid == 0 && P.seg_cap) P.seg_count[blockIdx.x] = 0; return; }

    for (uint32_t i = tid; i < sizeof(GenericCfg) / 4; i += NT)
        reinterpret_cast<uint32_t*>(&s_G)[i] = reinterpret_cast<const uint32_t*>(Gp)[i];
    if (P.hist_smem) for (uint32_t i = tid; i < T.n_keys; i += NT) hist[i] = 0;
    for (uint32_t i = tid; i < G_::ROWS; i += NT) rowmask[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&s_bar, 1); s_qn = 0; }
    const DevCfg& C = s_G.c;

    const uint32_t rows_loaded = NT + P.halo_rows;                     // owned + read-ahead rows of every tile
    const uint32_t load_bytes = rows_loaded * S;
    const uint64_t first_tile = beg / G_::OWN_BYTES;
    const uint64_t n_tiles = (end - 1) / G_::OWN_BYTES + 1;
    const uint8_t* __restrict__ buf = P.buf;
    QEntry* const myq = P.queue + (size_t)blockIdx.x * P.seg_cap;
    Acc acc{0, 0, 0, 0, 0, 0};
    unsigned long long gst[F2Q_N_STATS] = {0, 0, 0, 0, 0};            // stats of reads handled in place by the generic code
    uint32_t bar_phase = 0;

    // newline flags of one 16-byte chunk, bit i = byte i
    auto chunk_mask = [&](const uint8_t* p16) -> uint32_t {
        const uint4 v = *reinterpret_cast<const uint4*>(p16);
        const uint32_t z0 = eq_bytes(v.x, 0x0A0A0A0Au), z1 = eq_bytes(v.y, 0x0A0A0A0Au);
        const uint32_t z2 = eq_bytes(v.z, 0x0A0A0A0Au), z3 = eq_bytes(v.w, 0x0A0A0A0Au);
        const uint32_t lo = __dp4a(z0, 0x08040201u, __dp4a(z1, 0x80402010u, 0u));     // 128 * (flags of bytes 0..7)
        const uint32_t hi = __dp4a(z2, 0x08040201u, __dp4a(z3, 0x80402010u, 0u));     // 128 * (flags of bytes 8..15)
        return (lo >> 7) | (hi << 1);
    };

    for (;;) {
        __syncthreads();                                               // everyone is done with the previous tile
        // ---- ticket + load.  Thread 0 takes the ticket only now, so the tile's count is published ~one load later ----
        if (tid == 0) {
            const uint32_t tk = atomicAdd(P.ticket, 1u);
            s_tile = tk;
            const uint64_t b0 = (first_tile + tk) * G_::OWN_BYTES;
            if (first_tile + tk < n_tiles && b0 >= beg && b0 + load_bytes <= end) {       // interior tile: one TMA bulk copy
                mbar_expect_tx(&s_bar, load_bytes);
                tma_load_1d(tile, buf + b0, load_bytes, &s_bar);
            }
        }
        __syncthreads();
        const uint64_t t = first_tile + s_tile;
        if (t >= n_tiles) break;
        const uint64_t base = t * G_::OWN_BYTES;
        const bool interior = (base >= beg) && (base + load_bytes <= end);
        if (interior) {
            mbar_wait(&s_bar, bar_phase);
            bar_phase ^= 1;
        } else {
            // first / last tile of the range: loaded by the threads, bytes outside [beg, end) become 0
            for (uint32_t c = tid; c < load_bytes / 16; c += NT) {
                const uint64_t g = base + (uint64_t)c * 16;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (g + 16 > beg && g < end) {
                    v = ldg_stream(reinterpret_cast<const uint4*>(buf + g));
                    if (g < beg || g + 16 > end) {
                        uint32_t w[4] = {v.x, v.y, v.z, v.w};
                        #pragma unroll
                        for (int b = 0; b < 16; b++) {
                            const uint64_t pos = g + b;
                            if (pos < beg || pos >= end) w[b >> 2] &= ~(0xFFu << (8 * (b & 3)));
                        }
                        v = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                *reinterpret_cast<uint4*>(tile + c * 16) = v;
            }
            __syncthreads();
        }

        // ---- K1: newline mask of row tid; the read-ahead rows are scanned chunk-wise by the lanes of the last warp ----
        uint32_t m16[8];
        #pragma unroll
        for (int j = 0; j < 8; j++) m16[j] = 0;
        #pragma unroll
        for (int j = 0; j < CH; j++) m16[j] = chunk_mask(tile + tid * S + j * 16);
        const uint4 mv = make_uint4(m16[0] | (m16[1] << 16), m16[2] | (m16[3] << 16), m16[4] | (m16[5] << 16), m16[6] | (m16[7] << 16));
        rowmask[tid] = mv;
        Mask128 own;
        own.lo = (uint64_t)mv.x | ((uint64_t)mv.y << 32); own.hi = (uint64_t)mv.z | ((uint64_t)mv.w << 32);
        if (warp == NT / 32 - 1) {
            uint16_t* hm = reinterpret_cast<uint16_t*>(rowmask + NT);     // 8 u16 per row; slots >= CH stay 0
            for (uint32_t c = lane; c < P.halo_rows * CH; c += 32) {
                const uint32_t r = c / CH, j = c - r * CH;
                hm[r * 8 + j] = (uint16_t)chunk_mask(tile + (NT + r) * S + j * 16);
            }
        }
        const uint32_t cnt = __popc(mv.x) + __popc(mv.y) + __popc(mv.z) + __popc(mv.w);
        uint32_t incl = cnt;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
        #pragma unroll
        for (int w = 0; w < NT / 32; w++) { uint32_t x = s_wsum[w]; if (w < (int)warp) wbase += x; total += x; }
        const uint32_t excl = wbase + incl - cnt;                      // newlines of the tile before this row

        // ---- decoupled look-back for the line phase of the tile (warp 0) ----
        if (warp == 0) {
            const uint32_t A = total;
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
                if (t == n_tiles - 1 && !P.stitch) St->nl_total = (p0 + A) & LB_VALUE_MASK;
            }
        }
        __syncthreads();
        const uint32_t p0 = s_p0;
        // ---- reads whose header line ends in this row ----
        const uint32_t region_end = (uint32_t)min((uint64_t)load_bytes, end - base);       // valid bytes of the loaded region
        const bool region_has_eof = (base + load_bytes >= end);
        Mask128 m = own;
        uint32_t remaining = cnt;
        uint32_t skip = (4u - ((p0 + excl) & 3u)) & 3u;                // newlines of this row before the next header end
        while (skip < remaining) {
            for (uint32_t k = 0; k < skip; k++) m.pop_lowest();
            remaining -= skip + 1;
            skip = 3;
            const uint32_t s0 = tid * S + (uint32_t)m.pop_lowest() + 1u;          // first byte of the sequence line
            // the next three newlines, walking into the following rows if needed
            Mask128 cm = m; uint32_t crow = tid;
            uint32_t nl[3]; int have = 0;
            #pragma unroll
            for (int k = 0; k < 3; k++) {
                while (cm.empty() && crow + 1 < rows_loaded) {
                    crow++;
                    const uint4 r = rowmask[crow];
                    cm.lo = (uint64_t)r.x | ((uint64_t)r.y << 32); cm.hi = (uint64_t)r.z | ((uint64_t)r.w << 32);
                }
                if (cm.empty()) break;
                nl[k] = crow * S + (uint32_t)cm.pop_lowest();
                have = k + 1;
            }
            uint32_t e0 = 0, s3 = 0, e3 = 0;
            bool complete = false, spill = false;
            if (have == 3) {
                e0 = nl[0]; s3 = nl[1] + 1u; e3 = nl[2];
                complete = true;
                acc.last_end = max(acc.last_end, (unsigned long long)(base + e3 + 1));
            } else if (region_has_eof) {
                // no further newline exists: only an unterminated final quality line can complete the record
                if (eof && have == 2) {
                    e0 = nl[0]; s3 = nl[1] + 1u; e3 = region_end;
                    if (s3 < region_end) { complete = true; acc.last_end = max(acc.last_end, (unsigned long long)end); }
                }
            } else spill = true;

            if (spill) {
                // finish the geometry in global memory, then hand the read to the generic code
                uint64_t pos[4];
                pos[0] = base + s0 - 1;
                for (int k = 0; k < 3; k++) pos[k + 1] = (k < have) ? base + nl[k] : 0;
                uint64_t from = base + load_bytes;
                bool ok = true;
                for (int k = have + 1; k < 4; k++) {
                    uint64_t p = find_newline_global(buf, from, end);
                    if (p >= end) {
                        if (k == 3 && eof && from < end) pos[3] = end;             // unterminated final line
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
                    uint32_t slot = (POLICY == POLICY_GENERIC) ? 0xFFFFFFFFu : atomicAdd(&St->g_count, 1u);
                    if (slot < St->g_cap) P.gqueue[slot] = ge;
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
                    while (e0 > s0 && is_py_space(tile[e0 - 1])) e0--;
                    while (e3 > s3 && is_py_space(tile[e3 - 1])) e3--;
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
                            uint32_t slot = atomicAdd(&St->g_count, 1u);
                            if (slot < St->g_cap) P.gqueue[slot] = ge;
                            else g_process_read(s_G, T, E, O, (const uint8_t*)ge.seq_addr, (int)ge.seq_len, (const uint8_t*)ge.qual_addr, (int)ge.qual_len, gst);
                        } else if (C.miss <= 0) acc.nonal++;
                        else {
                            const uint32_t sl = atomicAdd(&s_qn, 1u);              // this CTA's private queue segment
                            if (sl < P.seg_cap) { QEntry e; e.key = key; e.bad = bad; e.len = klen; myq[sl] = e; }
                            else {                                                 // segment full: resolve right here
                                const uint32_t r = resolve_thread(T, C.miss, key, bad, klen);
                                if (r != RES_NONE) { atomicAdd(O.counts + r, 1ull); acc.imperfect++; } else acc.nonal++;
                            }
                        }
                    }
                }
            }
        }

    }

    // ---- CTA epilogue: queue segment length, histogram, statistics ----
    __syncthreads();
    if (tid == 0 && P.seg_cap) P.seg_count[blockIdx.x] = min(s_qn, P.seg_cap);
    if (P.hist_smem)
        for (uint32_t i = tid; i < T.n_keys; i += NT) { uint32_t v = hist[i]; if (v) atomicAdd(O.counts + i, (unsigned long long)v); }
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
    if (lane == 0 && le && !P.stitch) atomicMax(&St->last_rec_end, le);
}

}  // namespace f2q
