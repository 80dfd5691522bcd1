// tile.cuh — the fused hot kernel: ONE pass over the FASTQ bytes.
//
//   K1 line scan      per-row newline bitmasks (SWAR compare + dp4a gather) -> block scan of the counts -> a list of the
//                     tile's newline positions in shared memory; decoupled look-back across tiles for the line phase
//                     (records are "every 4 lines from byte 0", fast2q.py:324-328)
//   K2 extract        fixed-position window with Python slice clamping, rstrip, Phred fail-set test (fast2q.py:349-360)
//   K4 lookup/count   2-bit pack, exact probe of the packed-key table, shared-memory histogram (fast2q.py:365-367)
//   -> non-exact keys go to this CTA's segment of the resolver queue (K5), undecidable reads to the generic queue.
//
// Work decomposition.  A tile is NT rows of S = 16*CH bytes (CH odd: "thread t scans row t with 16-byte loads" is then
// bank-conflict free without a swizzle; the host picks CH so that S is just below the record length).  The first
// NT-H rows are OWNED by the tile, the last H rows are read-ahead (they are the first rows of the next tile): a read
// belongs to the tile that owns the newline ending its header line, and finds its other three newlines in the rows
// behind it.  Persistent CTAs of NT consumer threads + TWO control warps, three tile stages in shared memory, mbarriers:
//
//   loader warp, tile i     wait until stage i%3 is free (consumers arrived on `empty`) -> ticket (atomic, taken one tile
//                           ahead) -> TMA bulk copy (cp.async.bulk, completes on `full`)
//   look-back warp, tile j  wait for its newline count (`agg`) -> decoupled look-back over the status bytes of the tiles
//                           before it -> publish the inclusive prefix -> hand the line phase p0 to the consumers (`p0r`)
//                           All DRAM / L2 / atomic latency of the pipeline lives in these two warps.
//   consumers, iteration k  wait `full` of tile k+1 -> row masks, counts, ONE named barrier, position list, publish the
//                           aggregate (global status byte + `agg`)
//                           wait `p0r` of tile k (long done) -> parse its reads: thread q takes the q-th read of the tile
//                           (dense lanes, no row walk) -> arrive on `empty`
//
// A tile's aggregate never depends on a look-back (it is published right after the scan, one parse phase before the
// successors need it), tickets are taken in increasing order, and only look-back warps ever spin on other CTAs: the
// smallest unfinished tile can always make progress, so the look-back cannot deadlock.  Every wait is bounded: a
// protocol failure sets an error flag instead of hanging the device.
// The first / last tile of a range (bytes outside [beg, end) must read as 0) is loaded by the consumers instead of TMA.
#pragma once

#include "f2q_dev.cuh"
#include "generic.cuh"
#include "resolve.cuh"


namespace f2q {

// per-read code of the fused kernels: byte-wise generic | one fixed window, packed (fast1) | bit-parallel search sequences /
// several windows (flex.cuh): planes of 96 positions and <= 1 mismatch per search sequence | 160 positions and <= 3
enum { POLICY_GENERIC = 0, POLICY_FAST1 = 1, POLICY_FLEX_S = 2, POLICY_FLEX_B = 3 };
__host__ __device__ constexpr bool policy_is_flex(int p) { return p == POLICY_FLEX_S || p == POLICY_FLEX_B; }
__host__ __device__ constexpr int policy_pw(int p) { return p == POLICY_FLEX_B ? 5 : 3; }
__host__ __device__ constexpr int policy_k(int p) { return p == POLICY_FLEX_B ? 3 : 1; }

struct TileParams {
    const uint8_t* buf;        // 128-byte aligned base of the chunk buffer
    DevState* S;
    uint8_t* status;           // look-back status bytes, zeroed before the launch
    uint32_t* ticket;          // tile ticket counter, zeroed before the launch
    int stitch;                // 1: parse [0, S->stitch_len) of the carry buffer; 0: parse [S->beg, S->end)
    QEntry* queue;             // grid segments of seg_cap entries each
    uint32_t* seg_count;       // entries written per segment (one per CTA)
    uint32_t seg_cap;          // 0: resolve every non-exact key in place
    GEntry* gqueue;
    uint32_t hist_smem;        // 1: per-CTA shared-memory histogram of n_keys u32
    uint32_t halo_rows;        // H: read-ahead rows at the end of every tile (1 .. NT/2)
    uint32_t skip_if_spec_ok;  // 1: return at once when the speculative kernel's results were committed (spec.cuh)
    uint32_t debug;            // 1: count waits into DevState::dbg
};

// status byte of the decoupled look-back: bits [1:0] newline count mod 4, bits [3:2] 0 not ready / 1 aggregate / 2 prefix
constexpr uint32_t SB_AGG = 0x04u, SB_PREFIX = 0x08u;
// every device-side wait is bounded in TIME (clock64, about 4 s at 2 GHz).  A wait that runs out never traps (a trap would
// poison the process's whole CUDA context, and with it every other f2q_ctx on the GPU): it raises the CTA's abort flag,
// sets an error bit in DevState::error (f2q_end_sample then fails this ONE sample with F2Q_EINTERNAL) or, in the
// speculative kernel, fails the speculation so that the exact kernel redoes the chunk.  -DF2Q_TRAP_ON_TIMEOUT restores the trap.
constexpr long long WAIT_CYCLE_LIMIT = 1ll << 33;
#ifdef F2Q_TRAP_ON_TIMEOUT
#define F2Q_TIMEOUT_TRAP() __trap()
#else
#define F2Q_TIMEOUT_TRAP() ((void)0)
#endif
constexpr int TILE_STAGES = 3;
constexpr int TILE_CTRL_THREADS = 64;                // loader warp + look-back warp

template <int CH, int NT_>
struct TileGeom {
    static constexpr int S = CH * 16;                                   // bytes per row
    static constexpr int NT = NT_;                                      // rows per tile = consumer threads per CTA
    static constexpr int NS = TILE_STAGES;
    static constexpr int LOAD_BYTES = NT * S;
    static constexpr int STAGE_BYTES = ((LOAD_BYTES + 16 + 127) / 128) * 128;   // +16: word reads may run past the end
    static constexpr int CAP = 6 * NT;                                  // newline positions kept per tile
    static constexpr int NL_OFF = NS * STAGE_BYTES;                     // two u16[CAP + 8] position lists
    static constexpr int NL_STRIDE = (CAP + 8) * 2;
    static constexpr int EXCL_OFF = NL_OFF + 2 * NL_STRIDE;             // two u16[NT]: newlines of the tile before each row
    static constexpr int HIST_OFF = ((EXCL_OFF + 2 * NT * 2 + 15) / 16) * 16;
};

template <int CH, int NT>
__host__ __device__ inline size_t tile_smem_bytes(uint32_t hist_entries) {
    return (size_t)TileGeom<CH, NT>::HIST_OFF + (size_t)hist_entries * 4;
}

// ---- mbarrier / TMA bulk copy (sm_90+ PTX) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must not hang the GPU.  gerr = the sample's error word (DevState::error)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile uint32_t* abort, uint32_t* gerr, unsigned long long* dbg = nullptr) {
    if (*abort) return false;
    if (mbar_try_wait(bar, parity)) return true;                        // (the ordinary case reads no clock)
    const long long t0 = clock64();
    for (uint32_t spins = 0;; spins++) {
        if (mbar_try_wait(bar, parity)) { if (dbg && (threadIdx.x & 31) == 0) atomicAdd(dbg, (unsigned long long)(clock64() - t0)); return true; }
        if ((spins & 1023u) == 1023u) {
            if (*abort) return false;
            if (clock64() - t0 > WAIT_CYCLE_LIMIT) break;
        }
    }
    *abort = 1u;
    atomicOr(gerr, ERR_LOOKBACK_TIMEOUT);
    F2Q_TIMEOUT_TRAP();
    return false;
}
// global -> shared bulk copy by the TMA unit; completion is signalled on the mbarrier (bytes % 16 == 0, 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // earlier generic-proxy accesses to dst are ordered first
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int THREADS>
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

__device__ __forceinline__ uint32_t ld_volatile_u8(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u8(uint8_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u8 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// a + b issued as a multiply-add (a * one + b, `one` from reg_const(1)): integer adds and logic ops share the ALU pipe (one
// warp instruction per 2 cycles), multiply-adds run on the FMA pipe next to it; the byte-parallel loops are ALU-pipe bound
__device__ __forceinline__ uint32_t add_on_fma(uint32_t a, uint32_t one, uint32_t b) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}
// a constant the compiler must keep in a register (so that LOP3 can combine it with two other register operands)
__device__ __forceinline__ uint32_t reg_const(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// @region window_words
// W words starting at an arbitrary tile offset: W+1 aligned loads, W funnel shifts, everything unrolled so the
// independent words overlap in the pipeline
template <int W>
__device__ __forceinline__ void load_words(const uint8_t* tile, uint32_t o, uint32_t (&w)[W]) {
    const uint32_t a = o & ~3u, sh = (o & 3u) * 8u;
    uint32_t r[W + 1];
    #pragma unroll
    for (int i = 0; i <= W; i++) r[i] = *reinterpret_cast<const uint32_t*>(tile + a + 4 * i);
    #pragma unroll
    for (int i = 0; i < W; i++) w[i] = __funnelshift_r(r[i], r[i + 1], sh);
}

// quality test of a window of W words (n bytes, 4W-3 <= n <= 4W)
template <int W>
__device__ __forceinline__ bool qual_fails_w(const uint8_t* tile, uint32_t o, int n, uint32_t add_ge, uint32_t add_gt) {
    uint32_t w[W];
    load_words<W>(tile, o, w);
    if (n & 3) w[W - 1] &= (1u << (8 * (n & 3))) - 1u;                 // bytes past the slice become 0 (never fail)
    uint32_t acc = 0;
    #pragma unroll
    for (int i = 0; i < W; i++) {
        const uint32_t lo7 = w[i] & 0x7F7F7F7Fu;
        acc |= (lo7 + add_ge) & ~(lo7 + add_gt) & ~w[i];
    }
    return (acc & 0x80808080u) != 0;
}

// spread the low 16 bits to the even bit positions of a 32-bit word
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    return (x | (x << 1)) & 0x55555555u;
}

// 2-bit pack of a window of W words (n symbols).  bad = mask of symbols outside ACGT after upper(); their key bits are 0
template <int W>
__device__ __forceinline__ void pack_w(const uint8_t* tile, uint32_t o, int n, uint32_t& klo, uint32_t& khi, uint32_t& bad, bool act = true) {
    uint32_t w[W], d[W];
    load_words<W>(tile, o, w);
    if (n & 3) { const uint32_t keep = (1u << (8 * (n & 3))) - 1u; w[W - 1] = (w[W - 1] & keep) | (0x41414141u & ~keep); }   // pad with 'A'
    uint32_t lo = 0, hi = 0, any = 0;
    #pragma unroll
    for (int i = 0; i < W; i++) {
        const uint32_t codes = (w[i] >> 1) & 0x03030303u;
        uint32_t t = (codes | (codes >> 4)) & 0x00330033u;
        t = (t | (t >> 8)) & 0x3333u;                                  // nibble k = code of byte k
        d[i] = (w[i] & 0xDFDFDFDFu) ^ __byte_perm(0x47544341u, 0u, t);  // upper-cased byte vs the base its code stands for
        any |= d[i];
        const uint32_t p = (codes * 0x01041040u) >> 24;                 // 4 symbols -> 8 bits (one multiply, no carries)
        if (i < 4) lo |= p << (8 * (i & 3)); else hi |= p << (8 * (i & 3));
    }
    bad = 0;
    if (any && act) {                                                  // rare per read but common per warp: kept short (lanes without a read: skipped)
        #pragma unroll
        for (int i = 0; i < W; i++) {
            const uint32_t nz = (((d[i] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d[i]) & 0x80808080u;     // bit 7 of every non-zero byte
            bad |= ((nz * 0x00204081u) >> 28) << (4 * i);
        }
        lo &= ~(spread16(bad & 0xFFFFu) * 3u);
        hi &= ~(spread16(bad >> 16) * 3u);
    }
    klo = lo; khi = hi;
}

#define F2Q_WORDS_SWITCH(words, CALL)                                                                   \
    switch (words) {                                                                                    \
        case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break; \
        case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; default: CALL(8); break; \
    }

// @region slow_path
__device__ __noinline__ uint64_t find_newline_global(const uint8_t* buf, uint64_t from, uint64_t end) {
    for (uint64_t p = from; p < end; p++) if (buf[p] == '\n') return p;
    return end;
}

// per-thread accumulators, reduced once per CTA
struct Acc {
    unsigned long long reads, perfect, imperfect, nonal, qfail, last_end;
};

// a read whose four lines do not all lie inside the loaded tile (or a tile with more newlines than the position list
// holds): finish the geometry in global memory and run the byte-wise generic code on it.  hdr_end = absolute buffer
// offset of the newline that ends the header line.  Rare by construction; exactness matters here, speed does not.
__device__ __noinline__ void slow_record(const uint8_t* buf, uint64_t hdr_end, uint64_t end, bool eof, const GenericCfg& G,
                                         const SlowArgs* X, Acc& acc, unsigned long long* gst) {
    uint64_t pos[4];
    pos[0] = hdr_end;
    uint64_t from = hdr_end + 1;
    for (int k = 1; k < 4; k++) {
        const uint64_t p = find_newline_global(buf, from, end);
        if (p >= end) {
            if (k == 3 && eof && from < end) { pos[3] = end; break; }      // unterminated final quality line (fast2q.py:324-328)
            return;                                                        // incomplete record: carried or dropped
        }
        pos[k] = p; from = p + 1;
    }
    acc.reads++;
    const unsigned long long le = (unsigned long long)(pos[3] + 1 < end ? pos[3] + 1 : end);
    acc.last_end = acc.last_end > le ? acc.last_end : le;
    const uint8_t* Rp = buf + pos[0] + 1; const uint8_t* Qp = buf + pos[2] + 1;
    g_process_read(G, X->T, X->E, X->O, Rp, g_rstrip(Rp, (int)(pos[1] - pos[0] - 1)), Qp, g_rstrip(Qp, (int)(pos[3] - pos[2] - 1)), gst);
}

// the same geometry, but the read is only QUEUED for the generic kernel (k_generic_queue runs after the speculation was
// verified): the flex policies use it, because an Extract+Count insert done here could not be taken back
__device__ __noinline__ void slow_record_defer(const uint8_t* buf, uint64_t hdr_end, uint64_t end, bool eof, GEntry* gqueue, DevState* St, Acc& acc) {
    uint64_t pos[4];
    pos[0] = hdr_end;
    uint64_t from = hdr_end + 1;
    for (int k = 1; k < 4; k++) {
        const uint64_t p = find_newline_global(buf, from, end);
        if (p >= end) {
            if (k == 3 && eof && from < end) { pos[3] = end; break; }
            return;
        }
        pos[k] = p; from = p + 1;
    }
    acc.reads++;
    const unsigned long long le = (unsigned long long)(pos[3] + 1 < end ? pos[3] + 1 : end);
    acc.last_end = acc.last_end > le ? acc.last_end : le;
    GEntry ge;
    ge.seq_addr = (uint64_t)(buf + pos[0] + 1); ge.seq_len = (uint32_t)(pos[1] - pos[0] - 1);
    ge.qual_addr = (uint64_t)(buf + pos[2] + 1); ge.qual_len = (uint32_t)(pos[3] - pos[2] - 1);
    const uint32_t slot = atomicAdd(&St->g_count, 1u);
    if (slot < St->g_cap) gqueue[slot] = ge;
    else St->spec_fail = 1u;
}

// @region fast1_read
// per-thread constants of the packed single-window policy (fixed position, one feature, <= 32 symbols)
struct Fast1Ctx {
    int c_start, c_end, c_fmax, c_miss;
    bool simple_slice;
    uint32_t add_ge, add_gt;           // SWAR constants of the Phred fail-set test: fail iff 33 <= byte <= c_fmax
    uint32_t* hist;                    // this CTA's shared-memory histogram, or nullptr (global atomics)
    QEntry* myq;                       // this CTA's segment of the non-exact key queue
    uint32_t seg_cap;
    uint32_t* s_qn;                    // shared-memory fill count of the segment
    GEntry* gqueue;
    DevState* St;
    const SlowArgs* X;                 // tables / outputs in global memory for the calls that are not inlined
    __device__ __forceinline__ void init(const GenericCfg* Gp) {
        c_start = Gp->c.starts[0]; c_end = c_start + Gp->c.length; c_fmax = Gp->c.fmax_ph; c_miss = Gp->c.miss;
        simple_slice = c_start >= 0 && Gp->c.length >= 0;
        add_ge = (0x80u - 33u) * 0x01010101u; add_gt = (0x80u - (uint32_t)(c_fmax + 1)) * 0x01010101u;
    }
};
struct Fast1Counts { uint32_t reads, perfect, imperfect, nonal, qfail; };

// one read of the packed policy: sequence line at tile[s0, e0), quality line at tile[s3, e3) (both before rstrip);
// gseq / gqual = global addresses of the same two lines (for reads deferred to the generic queue)
__device__ __forceinline__ void fast1_read(const Fast1Ctx& F, const uint8_t* tile, uint32_t s0, uint32_t e0, uint32_t s3, uint32_t e3,
                                           const uint8_t* gseq, const uint8_t* gqual, const GenericCfg& G, const LibTables& T,
                                           const EcTable& E, const Outputs& O, Fast1Counts& n, unsigned long long* gst, uint32_t lane) {
    // ---- K2: rstrip, window, Phred test ----
    if (is_py_space(tile[e0 - 1])) while (e0 > s0 && is_py_space(tile[e0 - 1])) e0--;     // (tile[s0 - 1] is the '\n' before the line)
    if (is_py_space(tile[e3 - 1])) while (e3 > s3 && is_py_space(tile[e3 - 1])) e3--;
    int lo, hi, qlo, qhi;
    if (F.simple_slice) {
        const int ls = (int)(e0 - s0), lq = (int)(e3 - s3);
        lo = min(F.c_start, ls); hi = min(F.c_end, ls); qlo = min(F.c_start, lq); qhi = min(F.c_end, lq);
    } else {
        py_slice((int)(e0 - s0), F.c_start, F.c_end, lo, hi);
        py_slice((int)(e3 - s3), F.c_start, F.c_end, qlo, qhi);
    }
    const int qn = qhi - qlo;
    if (F.c_fmax != 0 && qn > 0) {
        bool fails;
        #define F2Q_QCALL(W) fails = qual_fails_w<W>(tile, s3 + qlo, qn, F.add_ge, F.add_gt)
        F2Q_WORDS_SWITCH((qn + 3) >> 2, F2Q_QCALL)
        #undef F2Q_QCALL
        if (fails) { n.qfail++; return; }
    }
    // ---- K4: pack, exact lookup, count ----
    const uint32_t klen = (uint32_t)(hi - lo);
    uint32_t klo = 0, khi = 0, bad = 0;
    if (klen) {
        #define F2Q_PCALL(W) pack_w<W>(tile, s0 + lo, (int)klen, klo, khi, bad)
        F2Q_WORDS_SWITCH((klen + 3) >> 2, F2Q_PCALL)
        #undef F2Q_PCALL
    }
    const bool generic_len = (T.generic_len_mask >> min(klen, 63u)) & 1ull;
    uint32_t idx = SLOT_EMPTY;
    if (!generic_len && bad == 0) {
        if (T.cslots) { if (klen == T.c_len) idx = compact_lookup(T, klo, khi); }
        else idx = fast_lookup(T, ((uint64_t)khi << 32) | klo, klen);
    }
    if (idx != SLOT_EMPTY) {
        n.perfect++;
        if (F.hist) atomicAdd(F.hist + idx, 1u);
        else atomicAdd(O.counts + idx, 1ull);
        return;
    }
    if (generic_len) {
        // library keys of this length exist that the packed tables cannot hold
        GEntry ge; ge.seq_addr = (uint64_t)gseq; ge.seq_len = e0 - s0;
        ge.qual_addr = (uint64_t)gqual; ge.qual_len = e3 - s3;
        const uint32_t slot = atomicAdd(&F.St->g_count, 1u);
        if (slot < F.St->g_cap) F.gqueue[slot] = ge;
        else g_process_read(G, F.X->T, F.X->E, F.X->O, (const uint8_t*)ge.seq_addr, (int)ge.seq_len, (const uint8_t*)ge.qual_addr, (int)ge.qual_len, gst);
        return;
    }
    if (F.c_miss <= 0) { n.nonal++; return; }
    // non-exact key: one slot of this CTA's private queue segment, reserved once per warp
    const uint64_t key = ((uint64_t)khi << 32) | klo;
    const uint32_t peers = __activemask();
    const uint32_t leader = (uint32_t)__ffs((int)peers) - 1u;
    uint32_t sl = 0;
    if (lane == leader) sl = atomicAdd(F.s_qn, (uint32_t)__popc(peers));
    sl = __shfl_sync(peers, sl, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    if (sl < F.seg_cap) { QEntry e; e.key = key; e.bad = bad; e.len = klen; F.myq[sl] = e; }
    else {                                                     // segment full: resolve right here
        const uint32_t r = resolve_seed_thread(F.X->T, F.c_miss, key, bad, klen);
        if (r != RES_NONE) { atomicAdd(O.counts + r, 1ull); n.imperfect++; } else n.nonal++;
    }
}

// @region fast1_read_warp
// The same read, written for a CONVERGED warp (spec.cuh) and split in two so that the table lookup's L2 latency hides behind
// the scan of the next tile.  Every lane calls both halves, `valid` lanes hold a read; nothing returns early.
//   fast1_warp_issue   rstrip, window, Phred test, pack, reads for the generic queue; issues the two cuckoo loads
//   fast1_warp_commit  compares the loaded slots, counts, and reserves the queue slots of the warp's non-exact keys with one ballot
// Together they compute exactly what fast1_read computes.
struct Fast1Pending {
    uint2 ra, rb;              // the two candidate slots (cuckoo), or ra.x = feature index found by the probing tables
    uint32_t klo, khi, bad;
    uint32_t meta;             // [5:0] key length, bit 8: may be an exact hit, bit 9: counts as non-exact when it is not, bit 10: ra.x is the index
};

__device__ __forceinline__ Fast1Pending fast1_warp_issue(const Fast1Ctx& F, bool valid, const uint8_t* tile, uint32_t s0, uint32_t e0, uint32_t s3,
                                                         uint32_t e3, const uint8_t* gseq, const uint8_t* gqual, const GenericCfg& G,
                                                         const LibTables& T, const EcTable& E, const Outputs& O, Fast1Counts& n,
                                                         unsigned long long* gst) {
    if (!valid) { s0 = e0 = s3 = e3 = 0; }
    else {
        if (is_py_space(tile[e0 - 1])) while (e0 > s0 && is_py_space(tile[e0 - 1])) e0--;
        if (is_py_space(tile[e3 - 1])) while (e3 > s3 && is_py_space(tile[e3 - 1])) e3--;
    }
    int lo, hi, qlo, qhi;
    if (F.simple_slice) {
        const int ls = (int)(e0 - s0), lq = (int)(e3 - s3);
        lo = min(F.c_start, ls); hi = min(F.c_end, ls); qlo = min(F.c_start, lq); qhi = min(F.c_end, lq);
    } else {
        py_slice((int)(e0 - s0), F.c_start, F.c_end, lo, hi);
        py_slice((int)(e3 - s3), F.c_start, F.c_end, qlo, qhi);
    }
    const int qn = qhi - qlo;
    bool fails = false;
    if (F.c_fmax != 0 && qn > 0) {
        #define F2Q_QCALL(W) fails = qual_fails_w<W>(tile, s3 + qlo, qn, F.add_ge, F.add_gt)
        F2Q_WORDS_SWITCH((qn + 3) >> 2, F2Q_QCALL)
        #undef F2Q_QCALL
    }
    Fast1Pending p;
    const uint32_t klen = (uint32_t)(hi - lo);
    p.klo = 0; p.khi = 0; p.bad = 0;
    if (klen) {
        #define F2Q_PCALL(W) pack_w<W>(tile, s0 + lo, (int)klen, p.klo, p.khi, p.bad)
        F2Q_WORDS_SWITCH((klen + 3) >> 2, F2Q_PCALL)
        #undef F2Q_PCALL
    }
    n.qfail += (valid && fails) ? 1u : 0u;
    const bool live = valid && !fails;
    const bool generic_len = live && ((T.generic_len_mask >> min(klen, 63u)) & 1ull);
    if (__any_sync(0xffffffffu, generic_len)) {
        if (generic_len) {                                     // library keys of this length exist that the packed tables cannot hold
            GEntry ge; ge.seq_addr = (uint64_t)gseq; ge.seq_len = e0 - s0;
            ge.qual_addr = (uint64_t)gqual; ge.qual_len = e3 - s3;
            const uint32_t slot = atomicAdd(&F.St->g_count, 1u);
            if (slot < F.St->g_cap) F.gqueue[slot] = ge;
            else g_process_read(G, F.X->T, F.X->E, F.X->O, (const uint8_t*)ge.seq_addr, (int)ge.seq_len, (const uint8_t*)ge.qual_addr, (int)ge.qual_len, gst);
        }
    }
    bool can = live && !generic_len && p.bad == 0;
    p.ra = make_uint2(SLOT_EMPTY, 0); p.rb = make_uint2(0, 0);
    uint32_t direct = 0;
    if (T.cuckoo) {                                            // (uniform) every lane loads; lanes without a key read some slot and ignore it
        can = can && klen == T.c_len;
        uint32_t h1, h2;
        cuckoo_slots(T.ck_mul[0], T.ck_mul[1], T.ck_mul[2], T.ck_mul[3], T.ck_shift, p.klo, p.khi, h1, h2);
        p.ra = ldg_u2_here(reinterpret_cast<const uint2*>(T.cuckoo) + h1);
        p.rb = ldg_u2_here(reinterpret_cast<const uint2*>(T.cuckoo) + h2);
    } else {
        direct = 1u << 10;
        if (can) {
            if (T.cslots) { if (klen == T.c_len) p.ra.x = compact_lookup(T, p.klo, p.khi); }
            else p.ra.x = fast_lookup(T, ((uint64_t)p.khi << 32) | p.klo, klen);
        }
    }
    p.meta = klen | (can ? 1u << 8 : 0u) | ((live && !generic_len) ? 1u << 9 : 0u) | direct;
    return p;
}

// the ordinary read, W = words of the window known at compile time: both lines reach the end of the window (so the window
// is [c_start, c_end) exactly), the library is one-length with a cuckoo table and no key of that length outside it (the
// caller checked).  rstrip cannot change such a read unless a non-ACGT byte sits in the window AND the line ends in
// whitespace (a window of pure ACGT lies before any whitespace tail; whitespace is never in the Phred fail set): that
// case, short lines and everything else unusual take fast1_read.  Same results, a third of the instructions.
template <int W>
__device__ __forceinline__ Fast1Pending fast1_ord_issue(const Fast1Ctx& F, bool valid, const uint8_t* tile, uint32_t s0, uint32_t e0, uint32_t s3,
                                                        uint32_t e3, const uint8_t* gseq, const uint8_t* gqual, const GenericCfg& G,
                                                        const LibTables& T, const EcTable& E, const Outputs& O, Fast1Counts& n,
                                                        unsigned long long* gst, uint32_t lane) {
    const int c_len = F.c_end - F.c_start;
    bool ord = valid && (int)(e0 - s0) >= F.c_end && (int)(e3 - s3) >= F.c_end;
    const uint32_t so = ord ? s0 + (uint32_t)F.c_start : 0u, qo = ord ? s3 + (uint32_t)F.c_start : 0u;
    const bool fails = F.c_fmax != 0 && qual_fails_w<W>(tile, qo, c_len, F.add_ge, F.add_gt);
    Fast1Pending p;
    pack_w<W>(tile, so, c_len, p.klo, p.khi, p.bad, ord);
    if (ord && p.bad != 0 && is_py_space(tile[e0 - 1])) ord = false;
    if (valid && !ord) fast1_read(F, tile, s0, e0, s3, e3, gseq, gqual, G, T, E, O, n, gst, lane);
    n.qfail += (ord && fails) ? 1u : 0u;
    const bool live = ord && !fails;
    uint32_t h1, h2;
    cuckoo_slots(T.ck_mul[0], T.ck_mul[1], T.ck_mul[2], T.ck_mul[3], T.ck_shift, p.klo, p.khi, h1, h2);
    p.ra = ldg_u2_here(reinterpret_cast<const uint2*>(T.cuckoo) + h1);
    p.rb = ldg_u2_here(reinterpret_cast<const uint2*>(T.cuckoo) + h2);
    p.meta = (uint32_t)c_len | ((live && p.bad == 0) ? 1u << 8 : 0u) | (live ? 1u << 9 : 0u);
    return p;
}

__device__ __forceinline__ void fast1_warp_commit(const Fast1Ctx& F, const Fast1Pending& p, const LibTables& T, const Outputs& O, Fast1Counts& n,
                                                  uint32_t lane) {
    // value of the key: feature index << 1 | imperfect, the tie marker, or CK_NONE (see the cuckoo build in f2q_set_library)
    uint32_t v = CK_NONE;
    const bool direct = (p.meta & (1u << 10)) != 0;
    if (direct) { if (p.ra.x != SLOT_EMPTY) v = p.ra.x << 1; }
    else v = cuckoo_match(T, p.ra, p.rb, p.klo, p.khi);
    const bool can = (p.meta & (1u << 8)) != 0;
    const bool found = can && v != CK_NONE;
    const bool ambig = found && !direct && v == cuckoo_ambig(T);
    const bool hit = found && !ambig;
    const uint32_t idx = v >> 1, imperfect = v & 1u;
    n.perfect += (hit && !imperfect) ? 1u : 0u;
    n.imperfect += (hit && imperfect) ? 1u : 0u;
    if (hit) {
        if (F.hist) atomicAdd(F.hist + idx, 1u);
        else atomicAdd(O.counts + idx, 1ull);
    }
    const bool miss = (p.meta & (1u << 9)) && !hit;
    if (F.c_miss <= 0) { n.nonal += miss ? 1u : 0u; return; }
    // decided without the resolver: two or more keys at distance 1 (a tie for every m), or m = 1 and the complete
    // neighbour table does not know this pure-ACGT key of the library's length
    const bool final_nonal = miss && (ambig || (can && !direct && T.ck_neighbours != 0 && F.c_miss == 1));
    n.nonal += final_nonal ? 1u : 0u;
    const bool to_queue = miss && !final_nonal;
    const uint32_t mq = __ballot_sync(0xffffffffu, to_queue);
    if (mq) {
        uint32_t sl = 0;
        if (lane == 0) sl = atomicAdd(F.s_qn, (uint32_t)__popc(mq));
        sl = __shfl_sync(0xffffffffu, sl, 0) + (uint32_t)__popc(mq & ((1u << lane) - 1u));
        if (to_queue) {
            const uint64_t key = ((uint64_t)p.khi << 32) | p.klo;
            const uint32_t klen = p.meta & 63u;
            if (sl < F.seg_cap) { QEntry e; e.key = key; e.bad = p.bad; e.len = klen; F.myq[sl] = e; }
            else {                                                 // segment full: resolve right here
                const uint32_t r = resolve_seed_thread(F.X->T, F.c_miss, key, p.bad, klen);
                if (r != RES_NONE) { atomicAdd(O.counts + r, 1ull); n.imperfect++; } else n.nonal++;
            }
        }
    }
}

// @region lookback
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

// control warp: newlines (mod 4) of the range before tile `rel`, by decoupled look-back over the status bytes.
// Each lane inspects 16 tiles (one 16-byte load), the warp 512 tiles per step, nearest first; the walk stops at the nearest
// tile that already published an inclusive prefix and needs every tile between to have published its aggregate.
__device__ __noinline__ uint32_t tile_lookback(const uint8_t* status, uint64_t rel, uint32_t lane, volatile uint32_t* abort, uint32_t* gerr, unsigned long long* dbg) {
    const long long t_start = clock64();
    // (another CTA that gave up has set the sample's error word: stop waiting for its tiles)
    auto give_up = [&](uint32_t spins) -> bool {
        if (*abort) return true;
        if ((spins & 255u) != 255u) return false;
        if (clock64() - t_start <= WAIT_CYCLE_LIMIT && ld_volatile_u32(gerr) == 0u) return false;
        *abort = 1u; atomicOr(gerr, ERR_LOOKBACK_TIMEOUT); F2Q_TIMEOUT_TRAP();
        return true;
    };
    uint32_t p0 = 0;
    int64_t hi = (int64_t)rel;                                          // tiles [hi, rel) are accounted for
    if (rel > 0) {
        // first wait for the tile right before this one, with ONE lane polling ONE byte: tickets are taken in time order, so
        // when it has published, (nearly) everything before it has too.  Polling the wide window from the start would have
        // every look-back warp of the grid hammer the same few L2 lines, which delays the very stores they wait for.
        uint32_t b = 0, spins = 0;
        for (;;) {
            if (lane == 0) b = ld_volatile_u8(status + rel - 1);
            b = __shfl_sync(0xffffffffu, b, 0);
            if (b & 0x0Cu) break;
            if (give_up(++spins)) return 0;
            __nanosleep(200);
        }
        if (b & SB_PREFIX) return b & 3u;
    }
    while (hi > 0) {
        const int64_t q = ((hi - 1) >> 4) - (int64_t)lane;             // this lane's group: tiles 16q .. 16q+15
        uint32_t sum = 0, spins = 0;
        bool found = false, ok = false;
        for (;;) {
            if (!ok) {                                                  // (a group that was complete stays complete: only late lanes poll)
                uint4 v = make_uint4(0x08080808u, 0x08080808u, 0x08080808u, 0x08080808u);      // before the range: prefix 0
                if (q >= 0) v = ld_volatile_v4(status + 16 * q);
                uint32_t w[4] = {v.w, v.z, v.y, v.x};                   // nearest word first
                if (lane == 0 && (hi & 15)) {                           // tiles >= hi are not part of this step: neutral (aggregate 0)
                    const int valid = (int)(hi & 15);
                    #pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int nb = min(max(valid - 4 * (3 - j), 0), 4);     // valid bytes of word j (word 3-j in memory order)
                        const uint32_t keep = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
                        w[j] = (w[j] & keep) | (0x04040404u & ~keep);
                    }
                }
                found = false; ok = true; sum = 0;
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (found) continue;
                    const uint32_t pre = (w[j] >> 3) & 0x01010101u, rdy = ((w[j] >> 2) | (w[j] >> 3)) & 0x01010101u;
                    uint32_t vals = w[j] & 0x03030303u;
                    if (pre) {
                        const uint32_t pi = (31u - (uint32_t)__clz((int)pre)) >> 3;          // nearest prefix byte of this word
                        const uint32_t need = pi >= 3 ? 0u : (0x01010101u << (8 * (pi + 1)));
                        ok = ok && ((rdy & need) == need);
                        vals &= 0xFFFFFFFFu << (8 * pi);
                        found = true;
                    } else ok = ok && (rdy == 0x01010101u);
                    sum += (vals * 0x01010101u) >> 24;
                }
            }
            const uint32_t has = __ballot_sync(0xffffffffu, found);
            const uint32_t first = has ? (uint32_t)__ffs((int)has) - 1u : 32u;
            const bool wait = !ok && lane <= first;
            if (!__any_sync(0xffffffffu, wait)) {
                p0 += __reduce_add_sync(0xffffffffu, lane <= first ? sum : 0u);
                if (has) return p0 & 3u;
                break;
            }
            if (lane > first) ok = true;                                // beyond the nearest prefix: never needed again in this step
            if (dbg && lane == 0) atomicAdd(dbg + 5, 1ull);
            if (give_up(++spins)) return 0;
            __nanosleep(400);                                           // (a tile that published out of order: rare)
        }
        hi = (((hi - 1) >> 4) - 31) << 4;                               // everything from this lane-31 group upwards is summed
    }
    return p0 & 3u;
}

// @region kernel_prologue
template <int POLICY, int CH, int NT>
__global__ void __launch_bounds__(NT + TILE_CTRL_THREADS, POLICY == POLICY_FAST1 ? (512 / NT) : (256 / NT))
k_tile(TileParams P, const GenericCfg* __restrict__ Gp, LibTables T, EcTable E, Outputs O, const SlowArgs* __restrict__ X) {
    using G_ = TileGeom<CH, NT>;
    constexpr int S = G_::S, NS = G_::NS, CAP = G_::CAP;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + G_::HIST_OFF);
    __shared__ uint32_t s_wsum[2][NT / 32];
    __shared__ uint32_t s_ticket[NS], s_mode[NS];                      // per stage: tile ticket, 0 none / 1 TMA / 2 loaded by the consumers
    __shared__ uint32_t s_total_own[NS], s_total_all[NS], s_p0[NS];
    __shared__ uint32_t s_qn, s_abort;
    __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS], bar_agg[NS], bar_p0[NS], bar_go[NS];
    __shared__ GenericCfg s_G;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevState* St = P.S;
    const uint64_t beg = P.stitch ? 0ull : St->beg;
    const uint64_t end = P.stitch ? (uint64_t)St->stitch_len : St->end;
    const bool eof = P.stitch ? (St->stitch_eof != 0) : (St->is_last != 0);
    if (P.skip_if_spec_ok && St->spec_ok) return;
    if (end <= beg) { if (tid == 0 && P.seg_cap) P.seg_count[blockIdx.x] = 0; return; }

    if (POLICY == POLICY_GENERIC)                                      // the packed policy keeps its few scalars in registers
        for (uint32_t i = tid; i < sizeof(GenericCfg) / 4; i += blockDim.x)
            reinterpret_cast<uint32_t*>(&s_G)[i] = reinterpret_cast<const uint32_t*>(Gp)[i];
    if (P.hist_smem) for (uint32_t i = tid; i < T.n_keys; i += blockDim.x) hist[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], NT / 32); mbar_init(&bar_agg[s], 1); mbar_init(&bar_p0[s], 1); mbar_init(&bar_go[s], 1);
            s_mode[s] = 0;
        }
        mbar_fence_init();
        s_qn = 0; s_abort = 0;
    }
    __syncthreads();
    volatile uint32_t* abort = &s_abort;
    uint32_t* const gerr = &St->error;
    unsigned long long* dbg = P.debug ? St->dbg : nullptr;

    const uint32_t own_rows = NT - P.halo_rows;
    const uint32_t own_bytes = own_rows * S;
    const uint64_t first_tile = beg / own_bytes;
    const uint64_t n_tiles = (end - 1) / own_bytes + 1;
    const uint8_t* __restrict__ buf = P.buf;

    // =====================================================================================================
    // @region control_warps
    if (warp == NT / 32) {
        // ---- loader warp: stage free -> ticket -> bulk copy.  It never waits for anything but the consumers, so a tile's
        // bytes (and with them its aggregate) never depend on a look-back.  A CTA must not own CONSECUTIVE tiles: its
        // second would be scanned only after its first is parsed, and the CTA owning the tile behind would queue up
        // behind that — a convoy through all CTAs.  So the tickets of the first tiles are spaced by a load latency
        // (by then every other CTA has taken its own), and later ones are taken only when a stage is free.
        for (uint32_t i = 0;; i++) {
            const uint32_t s = i % NS;
            if (i >= (uint32_t)NS) { if (!mbar_wait(&bar_empty[s], ((i / NS) - 1u) & 1u, abort, gerr, dbg ? dbg + 0 : nullptr)) break; }   // tile i-NS is parsed
            else if (i >= 1) { if (!mbar_wait(&bar_full[i - 1], 0u, abort, gerr)) break; }
            uint32_t tk = 0;
            if (lane == 0) tk = atomicAdd(P.ticket, 1u);
            tk = __shfl_sync(0xffffffffu, tk, 0);
            const uint64_t t = first_tile + tk;
            uint32_t mode = 0;
            if (t < n_tiles) {
                const uint64_t b0 = t * own_bytes;
                mode = (b0 >= beg && b0 + G_::LOAD_BYTES <= end) ? 1u : 2u;
            }
            if (lane == 0) {
                s_ticket[s] = tk; s_mode[s] = mode;
                if (mode == 1) {                                        // interior tile: one TMA bulk copy
                    mbar_expect_tx(&bar_full[s], G_::LOAD_BYTES);
                    tma_load_1d(smem + s * G_::STAGE_BYTES, buf + t * own_bytes, G_::LOAD_BYTES, &bar_full[s]);
                } else mbar_arrive(&bar_full[s]);                       // the consumers load it themselves / nothing left
            }
            if (mode == 0) break;
        }
        return;
    }
    if (warp == NT / 32 + 1) {
        // ---- look-back warp: tile j's newline count (from the consumers' scan) -> walk the status bytes of the tiles
        // before it -> publish the inclusive prefix -> hand the line phase to the consumers.  The only code that ever
        // spins on other CTAs.
        for (uint32_t j = 0;; j++) {
            const uint32_t s = j % NS, par = (j / NS) & 1u;
            if (!mbar_wait(&bar_full[s], par, abort, gerr, dbg ? dbg + 7 : nullptr)) break;            // (only to learn whether tile j exists)
            if (s_mode[s] == 0) break;
            if (!mbar_wait(&bar_agg[s], par, abort, gerr, dbg ? dbg + 1 : nullptr)) break;
            // start the walk when the consumers begin to parse tile j-1: a whole parse phase before p0 is needed, and late
            // enough that the tiles before this one have (nearly always) published — polling early slows everything down
            if (!mbar_wait(&bar_go[s], par, abort, gerr)) break;
            const uint32_t A = s_total_own[s], tk = s_ticket[s];
            const long long tl0 = dbg ? clock64() : 0;
            const uint32_t p0 = P.debug == 2 ? 0u : tile_lookback(P.status, tk, lane, abort, gerr, dbg);
            if (dbg && lane == 0) atomicAdd(dbg + 4, (unsigned long long)(clock64() - tl0));
            if (lane == 0) {
                st_volatile_u8(P.status + tk, SB_PREFIX | ((p0 + A) & 3u));
                if (first_tile + tk == n_tiles - 1 && !P.stitch) St->nl_total = (p0 + A) & 3u;
                s_p0[s] = p0;
                mbar_arrive(&bar_p0[s]);
            }
        }
        return;
    }

    // =====================================================================================================
    // consumers
    const GenericCfg& G = (POLICY == POLICY_GENERIC) ? s_G : *Gp;
    Fast1Ctx F;
    F.init(Gp);
    F.hist = P.hist_smem ? hist : nullptr; F.myq = P.queue + (size_t)blockIdx.x * P.seg_cap; F.seg_cap = P.seg_cap; F.s_qn = &s_qn;
    F.gqueue = P.gqueue; F.St = St; F.X = X;
    Acc acc{0, 0, 0, 0, 0, 0};
    unsigned long long gst[F2Q_N_STATS] = {0, 0, 0, 0, 0};            // stats of reads handled by the generic code
    Fast1Counts cn{0, 0, 0, 0, 0};                                     // per-thread counts of this launch

    const uint32_t c0A = reg_const(0x0A0A0A0Au), c7F = reg_const(0x7F7F7F7Fu), c80 = reg_const(0x80808080u);
    // @region scan_chunk_mask
    // newline flags of one 16-byte chunk, bit i = byte i.  Exact per-byte compare in 3 operations per word:
    //   a = (w ^ 0x0A..) & 0x7F..; b = a + 0x7F..  (bit 7: low 7 bits differ); z = ~(b | w) & 0x80..  (w's own bit 7 must be 0)
    auto chunk_mask = [&](const uint8_t* p16) -> uint32_t {
        const uint4 v = *reinterpret_cast<const uint4*>(p16);
        const uint32_t z0 = ~(((v.x ^ c0A) & c7F) + c7F | v.x) & c80, z1 = ~(((v.y ^ c0A) & c7F) + c7F | v.y) & c80;
        const uint32_t z2 = ~(((v.z ^ c0A) & c7F) + c7F | v.z) & c80, z3 = ~(((v.w ^ c0A) & c7F) + c7F | v.w) & c80;
        const uint32_t lo = __dp4a(z0, 0x08040201u, __dp4a(z1, 0x80402010u, 0u));     // 128 * (flags of bytes 0..7)
        const uint32_t hi = __dp4a(z2, 0x08040201u, __dp4a(z3, 0x80402010u, 0u));     // 128 * (flags of bytes 8..15)
        return (lo >> 7) | (hi << 1);
    };

    // @region scan_wait_load
    // all consumers: newline masks, counts and position list of local tile i (its bytes are in stage i % NS or get
    // loaded here); ONE consumer barrier inside
    auto scan = [&](uint32_t i) {
        const uint32_t s = i % NS, par = i & 1u;
        uint8_t* tile = smem + s * G_::STAGE_BYTES;
        const uint64_t t = first_tile + s_ticket[s];
        const uint64_t base = t * own_bytes;
        if (s_mode[s] == 2) {
            // first / last tile of the range: loaded by the threads, bytes outside [beg, end) become 0
            for (uint32_t c = tid; c < G_::LOAD_BYTES / 16; c += NT) {
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
            consumer_barrier<NT>();
        }
        // @region scan_masks_blockscan
        uint32_t m16[8];
        #pragma unroll
        for (int j = 0; j < 8; j++) m16[j] = 0;
        #pragma unroll
        for (int j = 0; j < CH; j++) m16[j] = chunk_mask(tile + tid * S + j * 16);
        uint32_t mw[4] = {m16[0] | (m16[1] << 16), m16[2] | (m16[3] << 16), m16[4] | (m16[5] << 16), m16[6] | (m16[7] << 16)};
        const uint32_t cnt = __popc(mw[0]) + __popc(mw[1]) + __popc(mw[2]) + __popc(mw[3]);
        uint32_t incl = cnt;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) s_wsum[par][warp] = incl;
        consumer_barrier<NT>();                                        // (also: every consumer is done parsing tile i-2)
        uint32_t wv = lane < NT / 32 ? s_wsum[par][lane] : 0u, wi = wv;  // warp totals, scanned again inside every warp
        #pragma unroll
        for (int d = 1; d < NT / 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += y; }
        const uint32_t wbase = __shfl_sync(0xffffffffu, wi - wv, warp);
        const uint32_t excl = wbase + incl - cnt;                      // newlines of the tile before this row
        reinterpret_cast<uint16_t*>(smem + G_::EXCL_OFF)[par * NT + tid] = (uint16_t)min(excl, 0xFFFFu);
        if (tid == own_rows) {                                         // excl of the first read-ahead row = newlines of the owned rows
            s_total_own[s] = excl;
            const uint64_t rel = t - first_tile;
            if (rel != 0) st_volatile_u8(P.status + rel, SB_AGG | (excl & 3u));
            mbar_arrive(&bar_agg[s]);
        }
        if (tid == NT - 1) s_total_all[s] = excl + cnt;
        // @region scan_nlpos
        uint16_t* nl = reinterpret_cast<uint16_t*>(smem + G_::NL_OFF + par * G_::NL_STRIDE);
        uint32_t o = excl;
        const uint32_t rowbase = tid * S;
        #pragma unroll
        for (int w = 0; w < (CH + 1) / 2; w++) {
            uint32_t m = mw[w];
            #pragma unroll
            for (int step = 0; step < 2; step++) {                     // two predicated steps: no divergence for ordinary FASTQ
                const bool has = m != 0;
                const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
                if (has && o < (uint32_t)CAP) nl[o] = (uint16_t)(rowbase + 32u * w + b);
                o += has ? 1u : 0u;
                m &= m - 1u;
            }
            while (m) {                                                // three or more newlines within 32 bytes
                const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
                m &= m - 1u;
                if (o < (uint32_t)CAP) nl[o] = (uint16_t)(rowbase + 32u * w + b);
                o++;
            }
        }
    };

    // @region loop_top
    const long long t_begin = dbg ? clock64() : 0;
    bool live = mbar_wait(&bar_full[0], 0u, abort, gerr) && s_mode[0] != 0;
    if (live) { if (tid == 0) mbar_arrive(&bar_go[0]); scan(0); }
    for (uint32_t k = 0; live; k++) {
        const uint32_t s = k % NS, par = k & 1u, s1 = (k + 1) % NS;
        // ---- next tile first: its aggregate is public one parse phase before the successors need it ----
        if (!mbar_wait(&bar_full[s1], ((k + 1) / NS) & 1u, abort, gerr, dbg ? dbg + 2 : nullptr)) break;
        const bool more = s_mode[s1] != 0;
        if (more) scan(k + 1);
        else consumer_barrier<NT>();                                   // (the last scan's lists are visible to everyone)
        if (!mbar_wait(&bar_p0[s], (k / NS) & 1u, abort, gerr, dbg ? dbg + 3 : nullptr)) break;
        const uint32_t p0 = s_p0[s];
        if (tid == 0 && more) mbar_arrive(&bar_go[s1]);                // the look-back of tile k+1 may start now

        // @region parse_setup
        // ---- reads of tile k: thread q takes the q-th read whose header line ends in the owned rows ----
        const uint8_t* tile = smem + s * G_::STAGE_BYTES;
        const uint16_t* nl = reinterpret_cast<const uint16_t*>(smem + G_::NL_OFF + par * G_::NL_STRIDE);
        const uint64_t t = first_tile + s_ticket[s];
        const uint64_t base = t * own_bytes;
        const uint32_t total_own = s_total_own[s], total_all = s_total_all[s];
        const uint32_t jf = (4u - (p0 & 3u)) & 3u;                     // first newline of the tile that ends a header line

        if (total_all > (uint32_t)CAP) {
            // more newlines than the position list holds (a tile of very short lines): every owned row walks its own
            // header ends and finishes each read in global memory
            if (tid < own_rows) {
                uint32_t idx = reinterpret_cast<const uint16_t*>(smem + G_::EXCL_OFF)[par * NT + tid];   // exact: total_all <= NT*S < 65536
                for (uint32_t b = 0; b < (uint32_t)S; b++) {
                    if (tile[tid * S + b] != '\n') continue;
                    if (((p0 + idx) & 3u) == 0) slow_record(buf, base + tid * S + b, end, eof, G, X, acc, gst);
                    idx++;
                }
            }
        } else
        for (uint32_t j = jf + 4u * tid; j < total_own; j += 4u * NT) {
            const uint32_t h0 = nl[j];
            if (j + 3 >= total_all) {
                // the read's last newline is not in the loaded rows (a long record, or the end of the range where only
                // an unterminated final quality line can still complete it): finish it in global memory
                slow_record(buf, base + h0, end, eof, G, X, acc, gst);
                continue;
            }
            const uint32_t s0 = h0 + 1u;
            uint32_t e0 = nl[j + 1];
            const uint32_t s3 = (uint32_t)nl[j + 2] + 1u;
            uint32_t e3 = nl[j + 3];
            cn.reads++;
            acc.last_end = (unsigned long long)(base + e3 + 1);         // (a thread meets its reads in stream order)
            if (POLICY == POLICY_GENERIC) {
                const uint8_t* Rp = tile + s0; const uint8_t* Qp = tile + s3;
                g_process_read(G, X->T, X->E, X->O, Rp, g_rstrip(Rp, (int)(e0 - s0)), Qp, g_rstrip(Qp, (int)(e3 - s3)), gst);
                continue;
            }
            // @region parse_k2
            fast1_read(F, tile, s0, e0, s3, e3, buf + base + s0, buf + base + s3, G, T, E, O, cn, gst, lane);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[s]);                     // this warp is done with stage s
        live = more;
    }

    if (dbg && lane == 0) atomicAdd(dbg + 6, (unsigned long long)(clock64() - t_begin));
    // @region epilogue
    // ---- CTA epilogue (consumers): queue segment length, histogram, statistics ----
    consumer_barrier<NT>();
    if (tid == 0 && P.seg_cap) P.seg_count[blockIdx.x] = min(s_qn, P.seg_cap);
    if (P.hist_smem)
        for (uint32_t i = tid; i < T.n_keys; i += NT) { uint32_t v = hist[i]; if (v) atomicAdd(O.counts + i, (unsigned long long)v); }
    acc.reads += cn.reads; acc.perfect += cn.perfect; acc.imperfect += cn.imperfect; acc.nonal += cn.nonal; acc.qfail += cn.qfail;
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
