// spec.cuh — the streaming form of the fused hot kernel: warp-autonomous, speculative line phase, verified afterwards.
//
// Records are "every 4 lines from byte 0" (fast2q.py:324-328), so a byte range can only be parsed when the number of
// newlines before it (mod 4, the LINE PHASE) is known.  k_tile (tile.cuh) gets it exactly with a decoupled look-back
// across tiles; that chain couples every CTA to its predecessors and costs more than the parsing itself.  Here the chain
// is cut:
//
//   * the chunk is cut into RANGES of range_bytes; a warp takes a range by ticket and streams its tiles (32 rows of
//     S = 16*CH bytes + read-ahead rows) through its own 3-stage TMA ring in shared memory (cp.async.bulk + mbarrier,
//     SASS UBLKCP).  Inside a range the phase is carried in a register: no CTA barrier, no other warp is ever waited for.
//   * the phase at the START of a range (other than the first, which is a record start) is SPECULATED from the local
//     record structure: exactly one of the four alignments must show '+' on line 2, '@' on the next header and
//     len(line 1) == len(line 3) for the first two records.  Anything ambiguous fails the speculation.
//   * every range publishes (newline count mod 4, speculated phase).  k_spec_verify adds them up: if every speculated
//     phase equals the true prefix, the results (which were accumulated into a SCRATCH count vector and the CTA-private
//     queue segments) are committed by k_spec_merge; otherwise they are dropped and the exact kernel k_tile parses the
//     chunk (it is launched unconditionally and returns at once when the speculation held).  The outcome is therefore
//     always exactly the reference's, whatever the bytes are; only the speed depends on the FASTQ being ordinary.
//
// Per tile a warp does: wait for tile i+1 -> newline masks / positions of its 32 rows (lane = row) -> parse tile i
// (lane q = q-th read; reads running into tile i+1 use its first rows' positions and the read-ahead bytes) -> refill the
// stage of tile i.  Behind the last tile of a range the read-ahead rows in its own stage are scanned instead; the tiles
// of consecutive ranges form one sequence through the ring, so the next range's first tiles are already in flight.
#pragma once

#include <type_traits>

#include "tile.cuh"
#include "flex.cuh"

namespace f2q {

// stages of a warp's ring: three (scan tile i+1 and parse tile i while tile i+2 lands) up to 16 warps per CTA; with 20 or 24
// warps two stages — the other warps hide the exposed part of a warp's own load latency, and the shared memory fits
__host__ __device__ constexpr int spec_stages(int W) { return W > 16 ? 2 : 3; }
constexpr int SPEC_CAP = 6 * 32;                     // newline positions kept per tile (its 32 own rows)
constexpr uint32_t SPEC_MAX_HALO = 16;

struct SpecParams {
    const uint8_t* buf;        // 128-byte aligned base of the chunk buffer
    DevState* S;
    uint32_t* ticket;          // range ticket counter, zeroed before the launch
    uint8_t* rec;              // one byte per range, zeroed before the launch: 0x80 | speculated phase << 2 | newline count & 3
    uint64_t range_bytes;      // multiple of the tile's own bytes (32 * 16 * CH)
    QEntry* queue;             // grid segments of seg_cap entries each (fast1: QEntry, flex Counter: FlexQ, flex Extract+Count: u64 log)
    uint32_t* seg_count;
    uint32_t seg_cap;
    GEntry* gqueue;
    uint32_t hist_smem;        // 1: per-CTA shared-memory histogram of n_keys u32
    uint32_t halo_rows;        // H: read-ahead rows loaded behind the 32 own rows (1 .. SPEC_MAX_HALO)
};

// bytes behind the loaded rows of a stage that word reads may touch: the fast1 code reads a few words past a line's end, the
// flex code loads whole 32 * PW-byte lines (the bytes are masked by the line length afterwards, they only must be readable)
__host__ __device__ constexpr uint32_t spec_stage_pad(int policy) { return policy_is_flex(policy) ? (uint32_t)(32 * policy_pw(policy) + 16) : 16u; }

template <int CH>
struct SpecGeom {
    static constexpr int S = CH * 16;
    static constexpr int OWN = 32 * S;
    static constexpr int MW = (CH + 1) / 2;                            // 32-bit newline mask words per row
    static constexpr int NL_LIST = SPEC_CAP + 8;                       // u16 entries of one position list
    static constexpr int NL_HALO = 6 * (int)16 + 8;                    // ... of the list of a range's last read-ahead rows (SPEC_MAX_HALO rows)
    __host__ __device__ static constexpr uint32_t stage_bytes(uint32_t H, uint32_t pad) { return (((32u + H) * S + pad + 127u) / 128u) * 128u; }
    __host__ __device__ static constexpr uint32_t warp_bytes(uint32_t H, uint32_t pad, uint32_t ns) { return ((ns * stage_bytes(H, pad) + (2u * NL_LIST + NL_HALO) * 2u + 127u) / 128u) * 128u; }
};

template <int POLICY, int CH, int W>
__host__ __device__ inline size_t spec_smem_bytes(uint32_t H, uint32_t hist_entries) {
    return (size_t)W * SpecGeom<CH>::warp_bytes(H, spec_stage_pad(POLICY), spec_stages(W)) + (size_t)hist_entries * 4;
}

__device__ __forceinline__ uint64_t spec_n_ranges(uint64_t beg, uint64_t end, uint64_t own, uint64_t range_bytes, uint64_t& origin0) {
    origin0 = (beg / own) * own;
    const uint64_t n = (end - origin0) / range_bytes;                 // the last range takes the remainder
    return n ? n : 1;
}

// @region spec_kernel
template <int POLICY, int CH, int W>
__global__ void __launch_bounds__(W * 32, 1)
k_spec(SpecParams P, const GenericCfg* __restrict__ Gp, LibTables T, EcTable E, Outputs O, const SlowArgs* __restrict__ X,
       const __grid_constant__ FlexCfg FC) {
    using G_ = SpecGeom<CH>;
    constexpr int S = G_::S, OWN = G_::OWN, NS = spec_stages(W), CAP = SPEC_CAP, MW = G_::MW;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[W][NS];
    __shared__ uint32_t s_qn;
    __shared__ GenericCfg s_G;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DevState* St = P.S;
    const uint64_t beg = St->beg, end = St->end;
    const bool eof = St->is_last != 0;
    const uint32_t H = P.halo_rows;
    constexpr bool FLEX = policy_is_flex(POLICY);
    constexpr uint32_t PAD = spec_stage_pad(POLICY);
    const uint32_t stage_bytes = G_::stage_bytes(H, PAD), warp_bytes = G_::warp_bytes(H, PAD, NS), load_bytes = (32u + H) * S;
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + (size_t)W * warp_bytes);
    if (St->spec_off || end <= beg) return;                            // (a sample whose speculation failed once stays on the exact kernel)

    if (POLICY == POLICY_GENERIC || FLEX)
        for (uint32_t i = tid; i < sizeof(GenericCfg) / 4; i += blockDim.x)
            reinterpret_cast<uint32_t*>(&s_G)[i] = reinterpret_cast<const uint32_t*>(Gp)[i];
    if (P.hist_smem) for (uint32_t i = tid; i < T.n_keys; i += blockDim.x) hist[i] = 0;
    if (lane == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&bar_full[warp][s], 1);
        mbar_fence_init();
    }
    if (tid == 0) s_qn = 0;
    __syncthreads();

    const GenericCfg& G = (POLICY == POLICY_GENERIC || FLEX) ? s_G : *Gp;
    Fast1Ctx F;
    F.init(Gp);
    F.hist = P.hist_smem ? hist : nullptr; F.myq = P.queue + (size_t)blockIdx.x * P.seg_cap; F.seg_cap = P.seg_cap; F.s_qn = &s_qn;
    F.gqueue = P.gqueue; F.St = St; F.X = X;
    FlexCtx FX;
    FX.q = reinterpret_cast<FlexQ*>(P.queue); FX.log = reinterpret_cast<unsigned long long*>(P.queue);
    FX.q_cap = P.seg_cap;                                              // (flex: P.seg_cap is the capacity of the whole array)
    FX.hist = P.hist_smem ? hist : nullptr; FX.gqueue = P.gqueue; FX.St = St;
    FX.mode = Gp->c.mode; FX.miss = Gp->c.miss;
    FlexWarp FW{FX_NOBLOCK, FX_BLOCK};
    Acc acc{0, 0, 0, 0, 0, 0};
    unsigned long long gst[F2Q_N_STATS] = {0, 0, 0, 0, 0};
    Fast1Counts cn{0, 0, 0, 0, 0};
    // ordinary-read code (fast1_ord_issue): fixed window inside the line, one-length library in the cuckoo table
    uint32_t ord_words = 0;
    {
        const int c_len = F.c_end - F.c_start;
        if (F.simple_slice && c_len >= 1 && c_len <= 31 && T.cuckoo && T.c_len == (uint32_t)c_len && !((T.generic_len_mask >> c_len) & 1ull))
            ord_words = (uint32_t)(c_len + 3) >> 2;
    }

    uint64_t origin0;
    const uint64_t RB = P.range_bytes;
    const uint64_t n_ranges = spec_n_ranges(beg, end, OWN, RB, origin0);
    const uint8_t* __restrict__ buf = P.buf;
    uint8_t* const wsm = smem + (size_t)warp * warp_bytes;
    uint16_t* const nlist = reinterpret_cast<uint16_t*>(wsm + NS * stage_bytes);
    uint64_t* const bars = bar_full[warp];

    // @region spec_loader
    const uint32_t c0A = reg_const(0x0A0A0A0Au), c7F = reg_const(0x7F7F7F7Fu), c80 = reg_const(0x80808080u), one = reg_const(1u);
    auto chunk_mask = [&](const uint8_t* p16) -> uint32_t {            // newline flags of one 16-byte chunk (see tile.cuh)
        const uint4 v = *reinterpret_cast<const uint4*>(p16);
        const uint32_t z0 = ~(add_on_fma((v.x ^ c0A) & c7F, one, c7F) | v.x) & c80, z1 = ~(add_on_fma((v.y ^ c0A) & c7F, one, c7F) | v.y) & c80;
        const uint32_t z2 = ~(add_on_fma((v.z ^ c0A) & c7F, one, c7F) | v.z) & c80, z3 = ~(add_on_fma((v.w ^ c0A) & c7F, one, c7F) | v.w) & c80;
        const uint32_t lo = __dp4a(z0, 0x08040201u, __dp4a(z1, 0x80402010u, 0u));
        const uint32_t hi = __dp4a(z2, 0x08040201u, __dp4a(z3, 0x80402010u, 0u));
        return (lo >> 7) | (hi << 1);
    };

    // @region spec_scan
    // newline positions of 32 rows of the tile (lane scans row row0 + lane when act) -> nl[0 .. total), at most cap of them;
    // hcnt = newlines of the first H of these rows.  row0 = 0: the tile's own rows; row0 = 32: its read-ahead rows
    auto scan = [&](const uint8_t* tile, uint32_t row0, bool act, uint16_t* nl, uint32_t cap, uint32_t& total, uint32_t& hcnt) {
        uint32_t mw[MW];
        #pragma unroll
        for (int w = 0; w < MW; w++) mw[w] = 0;
        const uint32_t rowbase = (row0 + (act ? lane : 0u)) * S;
        #pragma unroll
        for (int j = 0; j < CH; j++) mw[j >> 1] |= chunk_mask(tile + rowbase + j * 16) << (16 * (j & 1));
        if (!act) {
            #pragma unroll
            for (int w = 0; w < MW; w++) mw[w] = 0;
        }
        uint32_t cnt = 0;
        #pragma unroll
        for (int w = 0; w < MW; w++) cnt += __popc(mw[w]);
        uint32_t incl = cnt;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        const uint32_t excl = incl - cnt;
        total = __shfl_sync(0xffffffffu, incl, 31);
        hcnt = __shfl_sync(0xffffffffu, excl, H);                      // (H <= 16 < 32)
        uint32_t o = excl;
        const bool fits = incl <= cap;                                 // (a tile with more newlines than that is not parsed at all)
        uint32_t crowded = 0;                                          // some 32-byte word of this row holds three or more newlines
        #pragma unroll
        for (int w = 0; w < MW; w++) {
            const uint32_t m = fits ? mw[w] : 0u;
            const uint32_t c = __popc(m), m1 = m & (m - 1u);
            // branch-free: a word without a first / second newline stores into the spare slot behind the list
            nl[c >= 1 ? o : cap + 7u] = (uint16_t)(rowbase + 32u * w + (uint32_t)__ffs((int)m) - 1u);
            nl[c >= 2 ? o + 1u : cap + 7u] = (uint16_t)(rowbase + 32u * w + (uint32_t)__ffs((int)m1) - 1u);
            crowded |= m1 & (m1 - 1u);
            o += c;
        }
        if (__any_sync(0xffffffffu, crowded != 0)) {                   // not ordinary FASTQ: the rest of those words, one branch per tile
            o = excl;
            #pragma unroll
            for (int w = 0; w < MW; w++) {
                const uint32_t m = fits ? mw[w] : 0u;
                uint32_t m2 = m & (m - 1u), k = o + 2;
                m2 &= m2 - 1u;
                while (m2) { nl[k++] = (uint16_t)(rowbase + 32u * w + (uint32_t)__ffs((int)m2) - 1u); m2 &= m2 - 1u; }
                o += __popc(m);
            }
        }
    };

    // @region spec_loop
    uint32_t phase = 0, range_cnt = 0, spec_p0 = 0, par_bits = 0;
    Fast1Pending pend;                                                 // meta == 0: nothing pending (committing it is a no-op)
    pend.ra = make_uint2(0, 0); pend.rb = make_uint2(0, 0); pend.klo = pend.khi = pend.bad = pend.meta = 0;
    uint16_t* const nl_halo = nlist + 2 * G_::NL_LIST;

    // @region spec_parse
    // ---- reads of one tile: lane q takes the q-th read whose header line ends in its own rows.  nlW = the tile's position
    // list (its first three read-ahead newlines get appended), nlB / hc / halo_off = where those come from
    auto parse_tile = [&](const uint8_t* ptile, uint64_t pbase, uint16_t* nlW, uint32_t total_own, const uint16_t* nlB, uint32_t hc,
                          uint32_t halo_off) {
        const uint32_t total_all = total_own + hc;
        if (total_own > (uint32_t)CAP || hc > (uint32_t)CAP) {
            if (lane == 0) St->spec_fail = 1u;                         // a tile of very short lines: left to the exact kernel
            return;
        }
        // a read needs at most three newlines behind the tile's own: append them, the list is then contiguous
        if (lane < 3u && lane < hc) nlW[total_own + lane] = (uint16_t)((uint32_t)nlB[lane] + halo_off);
        __syncwarp();
        const uint16_t* nlA = nlW;
        const uint32_t jf = (4u - (phase & 3u)) & 3u;                  // first newline of the tile that ends a header line
        auto passes = [&](auto Wc) {
            constexpr int WW = decltype(Wc)::value;                    // words of the window (ordinary-read code), 0 = general code
            for (uint32_t j0 = jf; j0 < total_own; j0 += 128u) {       // (uniform: the warp stays converged through a pass)
                const uint32_t j = j0 + 4u * lane;
                bool valid = j < total_own;
                if (valid && j + 3 >= total_all) {
                    // the read's last newline is not in the loaded rows (a long record, or the end of the chunk)
                    if constexpr (FLEX) slow_record_defer(buf, pbase + nlA[j], end, eof, P.gqueue, St, acc);
                    else slow_record(buf, pbase + nlA[j], end, eof, G, X, acc, gst);
                    valid = false;
                }
                uint32_t s0 = 0, e0 = 0, s3 = 0, e3 = 0;
                if (valid) {
                    s0 = (uint32_t)nlA[j] + 1u; e0 = nlA[j + 1]; s3 = (uint32_t)nlA[j + 2] + 1u; e3 = nlA[j + 3];
                    cn.reads++;
                    acc.last_end = (unsigned long long)(pbase + e3 + 1);
                }
                if (POLICY == POLICY_GENERIC) {
                    if (valid) {
                        const uint8_t* Rp = ptile + s0; const uint8_t* Qp = ptile + s3;
                        g_process_read(G, X->T, X->E, X->O, Rp, g_rstrip(Rp, (int)(e0 - s0)), Qp, g_rstrip(Qp, (int)(e3 - s3)), gst);
                    }
                } else if constexpr (FLEX) {
                    __syncwarp();
                    flex_read_warp<policy_pw(POLICY), policy_k(POLICY)>(FX, FC, valid, ptile, s0, e0, s3, e3, buf + pbase + s0, buf + pbase + s3, T, O, cn, FW, lane);
                } else {
                    __syncwarp();
                    if (j0 != jf) fast1_warp_commit(F, pend, T, O, cn, lane);          // (a second pass over the same tile: rare)
                    if constexpr (WW == 0) pend = fast1_warp_issue(F, valid, ptile, s0, e0, s3, e3, buf + pbase + s0, buf + pbase + s3, G, T, E, O, cn, gst);
                    else pend = fast1_ord_issue<WW>(F, valid, ptile, s0, e0, s3, e3, buf + pbase + s0, buf + pbase + s3, G, T, E, O, cn, gst, lane);
                }
            }
        };
        if (POLICY == POLICY_GENERIC || FLEX) passes(std::integral_constant<int, 0>{});
        else switch (ord_words) {
            case 4: passes(std::integral_constant<int, 4>{}); break;
            case 5: passes(std::integral_constant<int, 5>{}); break;
            case 6: passes(std::integral_constant<int, 6>{}); break;
            default: passes(std::integral_constant<int, 0>{}); break;
        }
    };

    // @region spec_ranges
    // One range after the other: the tiles of all ranges a warp takes form ONE sequence through the 3-stage ring (sequence
    // number mod 3 = stage), so the first tiles of the next range (its ticket is taken a range ahead) are already in flight
    // while the current one ends.  Step ti of a range: wait for tile ti and scan it (ti == nown: scan the read-ahead rows of
    // the last tile instead) -> count the lookups issued one step ago -> parse tile ti-1 -> refill its stage with the tile
    // three places further down the sequence.  No per-tile descriptor: everything is a function of (range, ti).
    struct RangeInfo { uint32_t r, nown, t_lo, t_hi; uint64_t rb; bool ok; };
    auto take_range = [&]() -> RangeInfo {
        RangeInfo q{0xFFFFFFFFu, 0, 0, 0, 0, false};
        uint32_t r = 0xFFFFFFFFu;
        if (lane == 0) { r = atomicAdd(P.ticket, 1u); if (ld_volatile_u32(&St->spec_fail)) r = 0xFFFFFFFFu; }
        r = __shfl_sync(0xffffffffu, r, 0);
        if ((uint64_t)r >= n_ranges) return q;
        q.r = r; q.ok = true;
        q.rb = origin0 + (uint64_t)r * RB;
        const uint64_t re = ((uint64_t)r == n_ranges - 1) ? end : q.rb + RB;
        q.nown = (uint32_t)((re - q.rb + OWN - 1) / OWN);
        // tiles [t_lo, t_hi) are plain TMA copies; the others touch bytes outside [beg, end) and are loaded by the lanes
        q.t_lo = q.rb >= beg ? 0u : 1u;
        q.t_hi = (end >= q.rb + load_bytes) ? (uint32_t)min((uint64_t)q.nown, (end - q.rb - load_bytes) / OWN + 1u) : 0u;
        return q;
    };
    auto issue_of = [&](const RangeInfo& q, uint32_t ti, uint32_t s) {
        if (q.ok && ti < q.nown && ti >= q.t_lo && ti < q.t_hi && lane == 0) {
            mbar_expect_tx(&bars[s], load_bytes);
            tma_load_1d(wsm + s * stage_bytes, buf + q.rb + (uint64_t)ti * OWN, load_bytes, &bars[s]);
        }
    };
    RangeInfo cur = take_range();
    constexpr uint32_t DEAD = 0x80000000u;                             // bit of par_bits: a bulk copy of this warp timed out
    uint32_t pre = 0, gs = 0;                                          // tiles of `cur` already issued; stage of its tile 0
    while (cur.ok) {
        const RangeInfo nxt = take_range();
        const uint32_t r = cur.r, nown = cur.nown, t_lo = cur.t_lo, t_hi = cur.t_hi;
        const uint64_t rb = cur.rb;
        uint32_t nxt_pre = 0;
        // place t of the sequence that starts at this range's tile 0 -> stage st
        auto issue_seq = [&](uint32_t t, uint32_t st) {
            if (t < nown) issue_of(cur, t, st);
            else if (t - nown < nxt.nown) { issue_of(nxt, t - nown, st); nxt_pre = t - nown + 1u; }
        };
        for (uint32_t t = pre; t < (uint32_t)NS; t++) issue_seq(t, (gs + t) % NS);
        range_cnt = 0; spec_p0 = 0; phase = 0;                         // (range 0 starts at a record start)
        uint32_t prev_total = 0, s = gs, sp = gs;                      // s = stage of tile ti, sp = stage of tile ti - 1
        for (uint32_t ti = 0;; ti++) {                                  // ti = 0 .. nown; left through `own` (a bound of its own, nown + 1, gets spilled)
            const uint32_t par = ti & 1u;
            uint16_t* const nl = nlist + par * G_::NL_LIST;
            uint32_t total = 0, hcnt = 0;
            const bool own = ti < nown;
            if (own) {
                uint8_t* const tile = wsm + s * stage_bytes;
                const uint64_t base = rb + (uint64_t)ti * OWN;
                if (ti >= t_lo && ti < t_hi) {
                    // (bounded in TIME: a copy that never lands fails the speculation — the exact kernel then redoes the chunk.
                    // Nothing traps and nothing leaves the loop early: the warp stops WAITING (DEAD), runs through the rest
                    // of its range on whatever the stage holds, takes no further range (take_range sees spec_fail), and all it
                    // counted is dropped with the failed speculation; see WAIT_CYCLE_LIMIT in tile.cuh.  The clock is read
                    // only after 1024 failed tries: never on the ordinary path)
                    if (!(par_bits & DEAD) && !mbar_try_wait(&bars[s], (par_bits >> s) & 1u)) {
                        long long tw = 0;
                        uint32_t spins = 0;
                        while (!mbar_try_wait(&bars[s], (par_bits >> s) & 1u)) {
                            if ((++spins & 1023u) != 0u) continue;
                            const long long now = clock64();
                            if (tw == 0) tw = now;
                            else if (now - tw > WAIT_CYCLE_LIMIT) { par_bits |= DEAD; if (lane == 0) St->spec_fail = 1u; F2Q_TIMEOUT_TRAP(); break; }
                        }
                    }
                    par_bits ^= 1u << s;
                } else {
                    // first / last tiles of the chunk: loaded by the lanes, bytes outside [beg, end) become 0
                    for (uint32_t c = lane; c < load_bytes / 16; c += 32) {
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
                    __syncwarp();
                }
                scan(tile, 0u, true, nl, (uint32_t)CAP, total, hcnt);
                __syncwarp();
                if (ti == 0 && r != 0) {
                    // @region spec_speculate
                    const uint32_t n = min(total, (uint32_t)CAP);
                    bool cand = false;
                    if (lane < 4) {
                        bool ok = true; int K = 0;
                        #pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const uint32_t j = lane + 4u * k;
                            if (j + 3 >= n) break;
                            const uint32_t a = nl[j], b = nl[j + 1], c = nl[j + 2], e = nl[j + 3];
                            ok = ok && tile[b + 1] == '+' && tile[e + 1] == '@' && (b - a) == (e - c);
                            K++;
                        }
                        cand = ok && K > 0;
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, cand) & 0xFu;
                    if (__popc(m) == 1) { spec_p0 = (4u - ((uint32_t)__ffs((int)m) - 1u)) & 3u; phase = spec_p0; }
                    else if (lane == 0) St->spec_fail = 1u;
                }
            } else {
                // behind the last tile of the range: the newlines of its read-ahead rows are found in its own stage
                uint32_t dummy;
                scan(wsm + sp * stage_bytes, 32u, lane < H, nl_halo, (uint32_t)G_::NL_HALO - 8u, hcnt, dummy);
                __syncwarp();
                if (hcnt > (uint32_t)G_::NL_HALO - 8u) hcnt = (uint32_t)CAP + 1u;      // too many to list
            }
            // the lookups issued one step ago have long arrived: count them
            if (POLICY == POLICY_FAST1) { fast1_warp_commit(F, pend, T, O, cn, lane); pend.meta = 0; }
            if (ti > 0) {
                parse_tile(wsm + sp * stage_bytes, rb + (uint64_t)(ti - 1) * OWN, nlist + (par ^ 1u) * G_::NL_LIST, prev_total,
                           own ? nl : nl_halo, hcnt, own ? (uint32_t)OWN : 0u);
                __syncwarp();
                phase += prev_total; range_cnt += prev_total;
                issue_seq(ti - 1 + NS, sp);                            // refill the stage of the parsed tile
            }
            if (!own) break;
            prev_total = total; sp = s; s = (s == (uint32_t)NS - 1u) ? 0u : s + 1u;
        }
        if (lane == 0) P.rec[r] = (uint8_t)(0x80u | (spec_p0 << 2) | (range_cnt & 3u));
        gs = s; pre = nxt_pre; cur = nxt;
    }
    if (POLICY == POLICY_FAST1) fast1_warp_commit(F, pend, T, O, cn, lane);
    if constexpr (FLEX) fx_block_fill(FX, FW, lane);

    // @region spec_epilogue
    __syncthreads();
    if (tid == 0 && P.seg_cap && !FLEX) P.seg_count[blockIdx.x] = min(s_qn, P.seg_cap);
    if (P.hist_smem)
        for (uint32_t i = tid; i < T.n_keys; i += blockDim.x) { const uint32_t v = hist[i]; if (v) atomicAdd(O.counts + i, (unsigned long long)v); }
    acc.reads += cn.reads; acc.perfect += cn.perfect; acc.imperfect += cn.imperfect; acc.nonal += cn.nonal; acc.qfail += cn.qfail;
    acc.perfect += gst[F2Q_STAT_PERFECT]; acc.imperfect += gst[F2Q_STAT_IMPERFECT];
    acc.nonal += gst[F2Q_STAT_NON_ALIGNED]; acc.qfail += gst[F2Q_STAT_QUALITY_FAILED];
    unsigned long long v[5] = {acc.reads, acc.perfect, acc.imperfect, acc.nonal, acc.qfail};
    #pragma unroll
    for (int k = 0; k < 5; k++) {
        unsigned long long x = v[k];
        #pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) x += __shfl_down_sync(0xffffffffu, x, dd);
        if (lane == 0 && x) atomicAdd(O.stats + k, x);
    }
    unsigned long long le = acc.last_end;
    #pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) le = max(le, __shfl_down_sync(0xffffffffu, le, dd));
    if (lane == 0 && le) atomicMax(&St->last_rec_end, le);
}

// @region spec_verify
// one CTA: true phase of every range = prefix sum of the published newline counts; the speculation holds iff every range
// finished and guessed exactly that.  On failure everything the speculative kernel left behind is reset.
constexpr int SPEC_VERIFY_THREADS = 1024;
__global__ void __launch_bounds__(SPEC_VERIFY_THREADS) k_spec_verify(DevState* St, const uint8_t* __restrict__ rec, uint64_t range_bytes,
                                                                    uint32_t own_bytes, uint32_t* seg_count, uint32_t n_segs,
                                                                    uint8_t* status, uint64_t n_status) {
    __shared__ uint32_t s_sum[SPEC_VERIFY_THREADS];
    __shared__ uint32_t s_bad, s_total;
    const uint32_t tid = threadIdx.x;
    const uint64_t beg = St->beg, end = St->end;
    if (tid == 0) { s_bad = (St->spec_off || St->spec_fail) ? 1u : 0u; s_total = 0; }
    __syncthreads();
    uint32_t total = 0;
    if (end > beg && !s_bad) {
        uint64_t origin0;
        const uint64_t n_ranges = spec_n_ranges(beg, end, own_bytes, range_bytes, origin0);
        const uint64_t per = (n_ranges + SPEC_VERIFY_THREADS - 1) / SPEC_VERIFY_THREADS;
        const uint64_t r0 = min(n_ranges, (uint64_t)tid * per), r1 = min(n_ranges, r0 + per);
        uint32_t sum = 0, bad = 0;
        for (uint64_t r = r0; r < r1; r++) { const uint32_t v = rec[r]; bad |= (v & 0x80u) ? 0u : 1u; sum += v & 3u; }
        s_sum[tid] = sum;
        __syncthreads();
        if (tid == 0) { uint32_t run = 0; for (int k = 0; k < SPEC_VERIFY_THREADS; k++) { const uint32_t x = s_sum[k]; s_sum[k] = run; run += x; } }
        __syncthreads();
        uint32_t ph = s_sum[tid];
        for (uint64_t r = r0; r < r1; r++) { const uint32_t v = rec[r]; if (r > 0 && ((v >> 2) & 3u) != (ph & 3u)) bad = 1u; ph += v & 3u; }
        if (bad) atomicOr(&s_bad, 1u);
        if (r1 == n_ranges && r0 < r1) s_total = ph;                   // the thread that owns the last range
    }
    __syncthreads();
    total = s_total;
    const bool ok = s_bad == 0;
    if (tid == 0) {
        St->spec_ok = ok ? 1u : 0u;
        if (ok) { if (end > beg) { St->nl_total = total & 3u; St->spec_commits++; } }
        else { St->last_rec_end = 0; St->g_count = 0; St->q_count = 0; St->spec_off = 1u; St->spec_fallbacks++; }
    }
    if (!ok) {
        for (uint32_t i = tid; i < n_segs; i += SPEC_VERIFY_THREADS) seg_count[i] = 0;
        // the exact kernel is about to parse the chunk: its look-back status bytes start from zero (no memset per chunk for
        // the ordinary case, in which that kernel returns at once)
        for (uint64_t i = tid; i < n_status; i += SPEC_VERIFY_THREADS) status[i] = 0;
    }
}

// commit (or drop) the scratch count vector of the speculative kernel; scratch is left zeroed for the next chunk
__global__ void __launch_bounds__(256) k_spec_merge(const DevState* St, unsigned long long* __restrict__ scratch,
                                                    unsigned long long* __restrict__ result, uint64_t n) {
    const bool ok = St->spec_ok != 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long v = scratch[i];
        if (v) { if (ok) result[i] += v; scratch[i] = 0; }
    }
}

}  // namespace f2q
