// flex_core.h — bit-parallel feature extraction for every configuration that is not "one fixed window":
// the delimiter modes of sequence_tinder (fast2q.py:215-285, border_finder :628-658, binary_subtract :601-626) and
// multi-feature keys (fast2q.py:332-363), as register-resident SWAR code for ONE read per thread.
//
// A read of up to 32*PW bases is turned ONCE into bit planes over its positions (bit p of a plane = position p):
//     b0, b1   the two bits of the base code (c >> 1) & 3       A=0 C=1 T=2 G=3
//     ok       the byte is exactly one of 'A','C','G','T'        (the delimiter search compares RAW bytes, fast2q.py:337)
//   quality line:
//     lq       the byte is in the Phred fail set [33, fmax]      (fast2q.py:1112-1129)
// From these:  M_c = positions whose byte is base c;  a search sequence d[0..s) with <= k mismatches is found at all
// positions at once by bit-sliced counting ("bitap" for Hamming distance): for every i the plane M_{d[i]} >> i tells where
// symbol i matches; k+1 thermometer planes (0 mismatches so far, <= 1, ...) are updated with one logic operation per
// word.  The reference's "first position >= start_place with p <= r - s" is the lowest set bit of the result inside a
// range mask.  Quality tests of a slice are "any lq bit in [a, b)", keys are cut out of b0/b1 with funnel shifts.
// That replaces s byte compares per offset and per search sequence by ~3 operations per symbol for ALL offsets.
//
// Everything here is a pure function over registers, compiled for the device by nvcc and for the host by g++
// (tests/native/hostcheck.cpp checks it against the oracle on CPU).  Search sequences must be pure ACGT and at most 32
// symbols, allowed mismatches at most FLEX_MAX_K; any other configuration takes the byte-wise generic path.
#pragma once

#include <stdint.h>

#include "synth_gen.h"      // F2Q_HD

namespace f2q {

constexpr int FLEX_MAX_K = 3;            // mismatches allowed in a search sequence on this path
constexpr int FLEX_MAX_PIECE = 32;       // symbols of one key piece on this path

#if defined(__CUDA_ARCH__)
#define F2Q_DP4A(a, b, c) __dp4a((unsigned)(a), (unsigned)(b), (unsigned)(c))
#define F2Q_FSHR(lo, hi, s) __funnelshift_r((lo), (hi), (s))
#define F2Q_POPC(x) __popc(x)
#define F2Q_FFS(x) __ffs((int)(x))
#else
inline uint32_t f2q_dp4a_host(uint32_t a, uint32_t b, uint32_t c) {
    for (int k = 0; k < 4; k++) c += ((a >> (8 * k)) & 0xFFu) * ((b >> (8 * k)) & 0xFFu);
    return c;
}
inline uint32_t f2q_fshr_host(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
#define F2Q_DP4A(a, b, c) f2q_dp4a_host((a), (b), (c))
#define F2Q_FSHR(lo, hi, s) f2q_fshr_host((lo), (hi), (s))
#define F2Q_POPC(x) __builtin_popcount(x)
#define F2Q_FFS(x) __builtin_ffs((int)(x))
#endif

// a + b issued as a multiply-add on the device (a * one + b, `one` an opaque register holding 1): the byte-parallel loops
// are bound by the integer ALU pipe (logic, shifts, adds: one warp instruction per 2 cycles); multiply-adds run on the
// FMA pipe next to it
#if defined(__CUDA_ARCH__)
F2Q_HD uint32_t flex_one() { uint32_t r; asm volatile("mov.u32 %0, 1;" : "=r"(r)); return r; }
F2Q_HD uint32_t flex_add(uint32_t a, uint32_t b, uint32_t one) { uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b)); return r; }
#else
inline uint32_t flex_one() { return 1u; }
inline uint32_t flex_add(uint32_t a, uint32_t b, uint32_t one) { return a * one + b; }
#endif

// search sequence, prepared by the host: pos[c] = bit i set iff symbol i has base code c
struct FlexDelim {
    uint32_t pos[4];
    int32_t len;             // 1 .. 32
    int32_t k;               // allowed mismatches, 0 .. FLEX_MAX_K
};

// NW = 8 * PW words starting at byte offset o of `base` (4-byte aligned base address): aligned loads + funnel shifts.
// Only the first 2 * NG words (NG groups of 8 bytes) are loaded; the others read as 0
template <int NW, int NG>
F2Q_HD void flex_load(const uint8_t* base, uint32_t o, uint32_t (&w)[NW]) {
    const uint32_t a = o & ~3u, sh = (o & 3u) * 8u;
    uint32_t r[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; i++) r[i] = (i <= 2 * NG) ? *reinterpret_cast<const uint32_t*>(base + a + 4 * i) : 0u;
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = (i < 2 * NG) ? F2Q_FSHR(r[i], r[i + 1], sh) : 0u;
}

// bits [0, left) of one word: 0 for left <= 0, all ones for left >= 32.  Device: the PTX shift clamps its amount at 32
// (1 << 32 = 0), so this is max, shift, subtract
F2Q_HD uint32_t flex_ones_below(int32_t left) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("{\n\t.reg .u32 t;\n\tmax.s32 t, %1, 0;\n\tshl.b32 t, 1, t;\n\tsub.u32 %0, t, 1;\n\t}" : "=r"(r) : "r"(left));
    return r;
#else
    return left >= 32 ? 0xFFFFFFFFu : left <= 0 ? 0u : ((1u << left) - 1u);
#endif
}

// bits [0, n) of a PW-word plane (n <= 32 * PW)
template <int PW>
F2Q_HD void flex_prefix_mask(uint32_t n, uint32_t (&m)[PW]) {
#pragma unroll
    for (int j = 0; j < PW; j++) m[j] = flex_ones_below((int32_t)n - 32 * j);
}

// 8 flag bits (bit k = byte k of the pair w0|w1 has `bit` set, scaled by `bit`'s value) placed at plane bit 8 * g
template <int G4>
F2Q_HD void flex_insert(uint32_t& word, uint32_t r, int scale_log2) {
    // r = mask8 << scale_log2; wanted: mask8 << (8 * G4)
    if (8 * G4 >= scale_log2) word |= r << (8 * G4 - scale_log2);
    else word |= r >> (scale_log2 - 8 * G4);
}

// sequence line -> code planes b0, b1 and the raw-validity plane ok (bits at positions >= len are 0 in ok; b0/b1 are
// don't-care there).  w = the line's bytes as NW = 8 * PW words
template <int PW, int NG>
F2Q_HD void flex_seq_planes(const uint32_t (&w)[8 * PW], uint32_t len, uint32_t (&b0)[PW], uint32_t (&b1)[PW], uint32_t (&ok)[PW]) {
#pragma unroll
    for (int j = 0; j < PW; j++) { b0[j] = 0; b1[j] = 0; ok[j] = 0; }
    const uint32_t one = flex_one();
#pragma unroll
    for (int g = 0; g < NG; g++) {                                     // (NG: the groups the longest line of the warp reaches)
        const uint32_t w0 = w[2 * g], w1 = w[2 * g + 1];
        const uint32_t r0 = F2Q_DP4A(w0 & 0x02020202u, 0x08040201u, F2Q_DP4A(w1 & 0x02020202u, 0x80402010u, 0u));     // 2 * flags
        const uint32_t r1 = F2Q_DP4A(w0 & 0x04040404u, 0x08040201u, F2Q_DP4A(w1 & 0x04040404u, 0x80402010u, 0u));     // 4 * flags
        // d == 0 per byte <=> the byte is 'A','C','G' or 'T': canonical byte of the code = 0x41 | (c & 6), and 'T' (bits 2:1 = 10)
        // is 0x54 = 0x45 ^ 0x11
        uint32_t nz[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t x = h ? w1 : w0;
            const uint32_t t = (x >> 2) & ~(x >> 1) & 0x01010101u;
            const uint32_t d = x ^ (0x41414141u | (x & 0x06060606u)) ^ (t * 0x11u);
            nz[h] = (flex_add(d & 0x7F7F7F7Fu, 0x7F7F7F7Fu, one) | d) & 0x80808080u;                                            // 0x80 per non-zero byte
        }
        const uint32_t rn = F2Q_DP4A(nz[0], 0x08040201u, F2Q_DP4A(nz[1], 0x80402010u, 0u));                           // 128 * flags
        switch (g & 3) {
            case 0: flex_insert<0>(b0[g >> 2], r0, 1); flex_insert<0>(b1[g >> 2], r1, 2); flex_insert<0>(ok[g >> 2], rn, 7); break;
            case 1: flex_insert<1>(b0[g >> 2], r0, 1); flex_insert<1>(b1[g >> 2], r1, 2); flex_insert<1>(ok[g >> 2], rn, 7); break;
            case 2: flex_insert<2>(b0[g >> 2], r0, 1); flex_insert<2>(b1[g >> 2], r1, 2); flex_insert<2>(ok[g >> 2], rn, 7); break;
            default: flex_insert<3>(b0[g >> 2], r0, 1); flex_insert<3>(b1[g >> 2], r1, 2); flex_insert<3>(ok[g >> 2], rn, 7); break;
        }
    }
    uint32_t lm[PW];
    flex_prefix_mask<PW>(len, lm);
#pragma unroll
    for (int j = 0; j < PW; j++) ok[j] = ~ok[j] & lm[j];               // (ok held the NOT-valid flags so far)
}

// quality line -> lq: bit p set iff 33 <= byte p <= fmax and p < len.  add_ge / add_gt as in Fast1Ctx (tile.cuh):
// (0x80 - 33) and (0x80 - (fmax + 1)) replicated; fmax == 0 (empty fail set) must be handled by the caller (lq = 0)
template <int PW, int NG>
F2Q_HD void flex_lowq_plane(const uint32_t (&w)[8 * PW], uint32_t len, uint32_t add_ge, uint32_t add_gt, uint32_t (&lq)[PW]) {
#pragma unroll
    for (int j = 0; j < PW; j++) lq[j] = 0;
    const uint32_t one = flex_one();
#pragma unroll
    for (int g = 0; g < NG; g++) {
        uint32_t f[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t x = w[2 * g + h], lo7 = x & 0x7F7F7F7Fu;
            f[h] = flex_add(lo7, add_ge, one) & ~flex_add(lo7, add_gt, one) & ~x & 0x80808080u;                                              // 0x80 per failing byte
        }
        const uint32_t r = F2Q_DP4A(f[0], 0x08040201u, F2Q_DP4A(f[1], 0x80402010u, 0u));
        switch (g & 3) {
            case 0: flex_insert<0>(lq[g >> 2], r, 7); break;
            case 1: flex_insert<1>(lq[g >> 2], r, 7); break;
            case 2: flex_insert<2>(lq[g >> 2], r, 7); break;
            default: flex_insert<3>(lq[g >> 2], r, 7); break;
        }
    }
    uint32_t lm[PW];
    flex_prefix_mask<PW>(len, lm);
#pragma unroll
    for (int j = 0; j < PW; j++) lq[j] &= lm[j];
}

// plane >> i (i in [0, 32 * PW))
template <int PW>
F2Q_HD void flex_shr(const uint32_t (&p)[PW], uint32_t i, uint32_t (&out)[PW]) {
    const uint32_t wsh = i >> 5, bsh = i & 31u;
#pragma unroll
    for (int j = 0; j < PW; j++) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int t = 0; t < PW; t++) {                                 // (select, no dynamic register indexing)
            if ((uint32_t)t == j + wsh) lo = p[t];
            if ((uint32_t)t == j + wsh + 1) hi = p[t];
        }
        out[j] = F2Q_FSHR(lo, hi, bsh);
    }
}

// search sequences are short (<= 32): plane >> i for i < 32 needs no word select
template <int PW>
F2Q_HD void flex_shr_small(const uint32_t (&p)[PW], uint32_t i, uint32_t (&out)[PW]) {
#pragma unroll
    for (int j = 0; j < PW; j++) out[j] = F2Q_FSHR(p[j], j + 1 < PW ? p[j + 1] : 0u, i);
}

// positions where the search sequence D matches with <= D.k mismatches: out bit p set iff Hamming(D, R[p : p + len)) <= k,
// counting every position (also those that overlap the end of the read: the caller restricts p to <= r - len).
// M[c] = positions whose byte is base c (0 elsewhere).  KK = the exact k of this instance, RW = result words wanted (the
// higher ones read as 0: no position up there can be returned)
template <int PW, int KK, int RW>
F2Q_HD void flex_search_k(const uint32_t (&M)[4][PW], const FlexDelim& D, uint32_t (&out)[PW]) {
    uint32_t th[KK + 1][RW];                                           // th[e] = at most e mismatches so far
#pragma unroll
    for (int e = 0; e <= KK; e++)
#pragma unroll
        for (int j = 0; j < RW; j++) th[e][j] = 0xFFFFFFFFu;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t todo = D.pos[c];
        while (todo) {                                                 // (uniform: the search sequence is a kernel parameter)
            const uint32_t i = (uint32_t)F2Q_FFS(todo) - 1u;
            todo &= todo - 1u;
            uint32_t m[RW];
#pragma unroll
            for (int j = 0; j < RW; j++) m[j] = F2Q_FSHR(M[c][j], j + 1 < PW ? M[c][j + 1] : 0u, i);
#pragma unroll
            for (int e = KK; e >= 1; e--)
#pragma unroll
                for (int j = 0; j < RW; j++) th[e][j] = (th[e][j] & m[j]) | th[e - 1][j];
#pragma unroll
            for (int j = 0; j < RW; j++) th[0][j] &= m[j];
        }
    }
#pragma unroll
    for (int j = 0; j < PW; j++) out[j] = j < RW ? th[KK][j] : 0u;
}

// K = compile-time bound on D.k; rw = result words the caller needs (uniform)
template <int PW, int K>
F2Q_HD void flex_search(const uint32_t (&M)[4][PW], const FlexDelim& D, uint32_t rw, uint32_t (&out)[PW]) {
    constexpr int LOW = PW > 1 ? PW - 1 : 1;
    const bool low = rw <= (uint32_t)LOW;
    if (D.k <= 0) { if (low) flex_search_k<PW, 0, LOW>(M, D, out); else flex_search_k<PW, 0, PW>(M, D, out); }
    else if (K >= 1 && D.k == 1) { if (low) flex_search_k<PW, (K >= 1 ? 1 : 0), LOW>(M, D, out); else flex_search_k<PW, (K >= 1 ? 1 : 0), PW>(M, D, out); }
    else if (K >= 2 && D.k == 2) { if (low) flex_search_k<PW, (K >= 2 ? 2 : 0), LOW>(M, D, out); else flex_search_k<PW, (K >= 2 ? 2 : 0), PW>(M, D, out); }
    else { if (low) flex_search_k<PW, K, LOW>(M, D, out); else flex_search_k<PW, K, PW>(M, D, out); }
}

// lowest set bit of plane p inside [from, to) (to <= 32 * PW), or -1
template <int PW>
F2Q_HD int flex_first(const uint32_t (&p)[PW], int from, int to) {
    if (from < 0) from = 0;
    if (to <= from) return -1;
    uint32_t lo[PW], hi[PW];
    flex_prefix_mask<PW>((uint32_t)from, lo);
    flex_prefix_mask<PW>((uint32_t)to, hi);
    int res = -1;
#pragma unroll
    for (int j = PW - 1; j >= 0; j--) {
        const uint32_t v = p[j] & hi[j] & ~lo[j];
        if (v) res = 32 * j + F2Q_FFS(v) - 1;
    }
    return res;
}

// any set bit of plane p inside the Python slice [a, b) of a line of n positions (the plane is 0 at positions >= n)
template <int PW>
F2Q_HD bool flex_any(const uint32_t (&p)[PW], int a, int b) {
    if (a < 0) a = 0;
    if (b > 32 * PW) b = 32 * PW;
    if (b <= a) return false;
    uint32_t lo[PW], hi[PW];
    flex_prefix_mask<PW>((uint32_t)a, lo);
    flex_prefix_mask<PW>((uint32_t)b, hi);
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < PW; j++) acc |= p[j] & hi[j] & ~lo[j];
    return acc != 0;
}

// spread the low 16 bits to the even bit positions
F2Q_HD uint32_t flex_spread16(uint32_t x) {
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    return (x | (x << 1)) & 0x55555555u;
}

// the key piece at positions [a, a + n) (n <= 32, a + n <= 32 * PW): 2-bit codes (symbol i at bits 2i, 2i+1; A=0 C=1 T=2 G=3)
// and notok = mask of its symbols whose byte is not exactly 'A','C','G','T' (their code bits are set to 0)
template <int PW>
F2Q_HD void flex_cut(const uint32_t (&b0)[PW], const uint32_t (&b1)[PW], const uint32_t (&ok)[PW], uint32_t a, uint32_t n,
                     uint64_t& codes, uint32_t& notok) {
    uint32_t s0[PW], s1[PW], so[PW];
    flex_shr<PW>(b0, a, s0);
    flex_shr<PW>(b1, a, s1);
    flex_shr<PW>(ok, a, so);
    const uint32_t keep = n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
    notok = ~so[0] & keep;
    const uint32_t v0 = s0[0] & keep & ~notok, v1 = s1[0] & keep & ~notok;
    const uint32_t lo = flex_spread16(v0 & 0xFFFFu) | (flex_spread16(v1 & 0xFFFFu) << 1);
    const uint32_t hi = flex_spread16(v0 >> 16) | (flex_spread16(v1 >> 16) << 1);
    codes = ((uint64_t)hi << 32) | lo;
}


// ------------------------------------------------------------------------------------------------------------
// one read: the key pieces of fast2q.py:332-363 (fixed windows or search sequences), from the planes
// ------------------------------------------------------------------------------------------------------------
constexpr int FLEX_ITER = 2;             // search iterations (--st / --us / --ds items) on this path

struct FlexCfg {
    int32_t n_iter, has_up, has_down, length;
    int32_t starts[FLEX_ITER];
    FlexDelim up[FLEX_ITER], down[FLEX_ITER];
    uint32_t fmax_ph, fmax_up, fmax_down;          // largest failing quality byte, 0 = empty fail set
    int32_t max_k;                                  // max over the search sequences' k
    int32_t eligible;                               // 1: this configuration can run on the bit-parallel path
};

struct FlexPiece {
    uint64_t codes;          // 2-bit codes of the piece's symbols, 0 where notok
    uint32_t notok;          // symbols whose RAW byte is not 'A','C','G','T' (lower case included: the caller looks at those bytes)
    uint32_t len, off;       // symbols; start position in the sequence line
};

// Python slice bounds seq[a:b] for a sequence of length n (as py_slice in f2q_dev.cuh)
F2Q_HD void flex_py_slice(int n, int a, int b, int& lo, int& hi) {
    if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
    if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
    if (b < a) b = a;
    lo = a; hi = b;
}

F2Q_HD void flex_lowq(const uint32_t* qw_dummy, uint32_t fmax, uint32_t& add_ge, uint32_t& add_gt) {
    (void)qw_dummy;
    add_ge = (0x80u - 33u) * 0x01010101u;
    add_gt = (0x80u - (fmax + 1u)) * 0x01010101u;
}

// Search-sequence modes.  sw / qw: the sequence / quality line as words (flex_load), r / q their lengths after rstrip (both
// <= 32 * PW), maxlen = the longest line of the warp (uniform), NG >= ceil(maxlen / 8) the 8-byte groups the planes are built for.
// Returns the number of pieces (>= 0; the key is their ':'-join), -1 when every iteration was flagged (quality_failed,
// fast2q.py:389-390), -(pieces + 1) <= -2 when a piece is longer than FLEX_MAX_PIECE (pc[] then holds every piece's length
// and offset, but codes only for the short ones).
template <int PW, int K, int NG>
F2Q_HD int flex_pieces_delim(const FlexCfg& C, const uint32_t (&sw)[8 * PW], uint32_t r, const uint32_t (&qw)[8 * PW], uint32_t q, uint32_t maxlen,
                             FlexPiece (&pc)[FLEX_ITER]) {
    uint32_t b0[PW], b1[PW], ok[PW], lq[PW], lqu[PW], lqd[PW];
    flex_seq_planes<PW, NG>(sw, r, b0, b1, ok);
    uint32_t age, agt;
    flex_lowq(nullptr, C.fmax_ph, age, agt);
    flex_lowq_plane<PW, NG>(qw, q, age, agt, lq);
    if (C.fmax_ph == 0) {
#pragma unroll
        for (int j = 0; j < PW; j++) lq[j] = 0;
    }
#pragma unroll
    for (int j = 0; j < PW; j++) { lqu[j] = lq[j]; lqd[j] = lq[j]; }
    if (C.has_up && C.fmax_up != C.fmax_ph) {                          // (uniform; the three thresholds are normally equal)
        flex_lowq(nullptr, C.fmax_up, age, agt);
        flex_lowq_plane<PW, NG>(qw, q, age, agt, lqu);
        if (C.fmax_up == 0) {
#pragma unroll
            for (int j = 0; j < PW; j++) lqu[j] = 0;
        }
    }
    if (C.has_down && C.fmax_down != C.fmax_ph) {
        flex_lowq(nullptr, C.fmax_down, age, agt);
        flex_lowq_plane<PW, NG>(qw, q, age, agt, lqd);
        if (C.fmax_down == 0) {
#pragma unroll
            for (int j = 0; j < PW; j++) lqd[j] = 0;
        }
    }
    uint32_t M[4][PW];
#pragma unroll
    for (int j = 0; j < PW; j++) {
        M[0][j] = ~b1[j] & ~b0[j] & ok[j]; M[1][j] = ~b1[j] & b0[j] & ok[j];
        M[2][j] = b1[j] & ~b0[j] & ok[j];  M[3][j] = b1[j] & b0[j] & ok[j];
    }
    int np = 0;
    bool any = false, slow = false;
#pragma unroll
    for (int i = 0; i < FLEX_ITER; i++) {
        if (i >= C.n_iter) break;
        int start, end;
        bool found = true;
        int u = 0, d = 0;
        const int ul = C.up[i].len, dl = C.down[i].len;
        if (C.has_up) {
            // a match starts at p <= longest line - ul: result words beyond that are never looked at
            uint32_t hit[PW];
            flex_search<PW, K>(M, C.up[i], maxlen >= (uint32_t)ul ? ((maxlen - (uint32_t)ul) >> 5) + 1u : 1u, hit);
            u = flex_first<PW>(hit, 0, (int)r - ul + 1);
            found = u >= 0;
        }
        if (C.has_down) {
            uint32_t hit[PW];
            flex_search<PW, K>(M, C.down[i], maxlen >= (uint32_t)dl ? ((maxlen - (uint32_t)dl) >> 5) + 1u : 1u, hit);
            d = flex_first<PW>(hit, C.has_up ? u + ul : 0, (int)r - dl + 1);
            found = found && d >= 0;
        }
        if (found && C.has_up && flex_any<PW>(lqu, u, u + ul)) found = false;
        if (found && C.has_down && flex_any<PW>(lqd, d, d + dl)) found = false;
        if (C.has_up && C.has_down) { start = u + ul; end = d; }
        else if (C.has_up) { start = u + ul; end = start + C.length; }
        else { start = d - C.length; end = d; }
        if (!found || end < start) continue;
        int lo, hi, qlo, qhi;
        flex_py_slice((int)r, start, end, lo, hi);
        flex_py_slice((int)q, start, end, qlo, qhi);
        if (flex_any<PW>(lq, qlo, qhi)) continue;
        any = true;
        const uint32_t n = (uint32_t)(hi - lo);
        FlexPiece p;
        p.codes = 0; p.notok = 0;
        if (n > (uint32_t)FLEX_MAX_PIECE) slow = true;                 // (its length is still reported: the caller may need no more)
        else flex_cut<PW>(b0, b1, ok, (uint32_t)lo, n, p.codes, p.notok);
        p.len = n; p.off = (uint32_t)lo;
        pc[np] = p;
        np++;
    }
    if (slow) return -(np + 1);
    return any ? np : -1;
}

// Fixed windows (--st a,b --l n, fast2q.py:349-360): no planes of the whole line, each window is loaded by itself.
// seq / qual: addresses such that the lines start at seq[so], qual[qo] (4-byte aligned base, see flex_load); a window is
// at most 32 symbols (flex_prepare checks C.length) = NGW <= 4 groups of 8.  Same return convention as flex_pieces_delim.
template <int NGW>
F2Q_HD int flex_pieces_fixed(const FlexCfg& C, const uint8_t* seq, uint32_t so, uint32_t r, const uint8_t* qual, uint32_t qo, uint32_t q,
                             FlexPiece (&pc)[FLEX_ITER]) {
    uint32_t age, agt;
    flex_lowq(nullptr, C.fmax_ph, age, agt);
    int np = 0;
    bool any = false;
#pragma unroll
    for (int i = 0; i < FLEX_ITER; i++) {
        if (i >= C.n_iter) break;
        const int start = C.starts[i], end = start + C.length;
        int lo, hi, qlo, qhi;
        flex_py_slice((int)r, start, end, lo, hi);
        flex_py_slice((int)q, start, end, qlo, qhi);
        uint32_t w[8], lq[1] = {0u};
        if (C.fmax_ph != 0) {
            flex_load<8, NGW>(qual, qo + (uint32_t)qlo, w);
            flex_lowq_plane<1, NGW>(w, (uint32_t)(qhi - qlo), age, agt, lq);
        }
        flex_load<8, NGW>(seq, so + (uint32_t)lo, w);
        uint32_t b0[1], b1[1], ok[1];
        const uint32_t n = (uint32_t)(hi - lo);
        flex_seq_planes<1, NGW>(w, n, b0, b1, ok);
        if (lq[0]) continue;                                           // the window's quality slice holds a failing byte
        any = true;
        const uint32_t keep = flex_ones_below((int32_t)n);
        FlexPiece p;
        p.notok = ~ok[0] & keep;
        const uint32_t v0 = b0[0] & keep & ~p.notok, v1 = b1[0] & keep & ~p.notok;
        const uint32_t clo = flex_spread16(v0 & 0xFFFFu) | (flex_spread16(v1 & 0xFFFFu) << 1);
        const uint32_t chi = flex_spread16(v0 >> 16) | (flex_spread16(v1 >> 16) << 1);
        p.codes = ((uint64_t)chi << 32) | clo;
        p.len = n; p.off = (uint32_t)lo;
        pc[np] = p;
        np++;
    }
    return any ? np : -1;
}

// one read: fixed windows or search sequences.  seq / qual: 4-byte aligned base addresses, the lines start at seq[so] and
// qual[qo]; maxlen >= max(r, q) is uniform over the warp
template <int PW, int K>
F2Q_HD int flex_pieces(const FlexCfg& C, const uint8_t* seq, uint32_t so, uint32_t r, const uint8_t* qual, uint32_t qo, uint32_t q, uint32_t maxlen,
                       FlexPiece (&pc)[FLEX_ITER]) {
    if (!(C.has_up || C.has_down)) {
        // (uniform) a window of <= 24 symbols needs three 8-byte groups, longer ones four
        if (C.length <= 24) return flex_pieces_fixed<3>(C, seq, so, r, qual, qo, q, pc);
        return flex_pieces_fixed<4>(C, seq, so, r, qual, qo, q, pc);
    }
    uint32_t sw[8 * PW], qw[8 * PW];
    constexpr int NGS = 4 * PW - 2;                                    // (75 bp reads: 10 groups of the 96-position planes)
    if (maxlen <= 8u * NGS) {
        flex_load<8 * PW, NGS>(seq, so, sw);
        flex_load<8 * PW, NGS>(qual, qo, qw);
        return flex_pieces_delim<PW, K, NGS>(C, sw, r, qw, q, maxlen, pc);
    }
    flex_load<8 * PW, 4 * PW>(seq, so, sw);
    flex_load<8 * PW, 4 * PW>(qual, qo, qw);
    return flex_pieces_delim<PW, K, 4 * PW>(C, sw, r, qw, q, maxlen, pc);
}

// host side: can this configuration run on the bit-parallel path?  (search sequences pure ACGT, 1..32 symbols, <= FLEX_MAX_K
// mismatches, at most FLEX_ITER iterations, feature length 0..32)
inline bool flex_prepare(const f2q_config& g, FlexCfg& C) {
    C = FlexCfg{};
    C.n_iter = g.n_iter; C.has_up = g.has_up != 0; C.has_down = g.has_down != 0; C.length = g.length;
    auto fmax = [](int ph) { if (ph <= 0) ph = 1; int n = ph - 1 < 94 ? ph - 1 : 94; return (uint32_t)(n == 0 ? 0 : 33 + n - 1); };
    C.fmax_ph = fmax(g.phred); C.fmax_up = fmax(g.qual_up); C.fmax_down = fmax(g.qual_down);
    if (g.n_iter < 1 || g.n_iter > FLEX_ITER) return false;
    const bool delim = C.has_up || C.has_down;
    if (!(C.has_up && C.has_down) && (g.length < 0 || g.length > FLEX_MAX_PIECE)) return false;
    auto prep = [](const uint8_t* s, int len, int k, FlexDelim& D) {
        if (len < 1 || len > 32 || k < 0 || k > FLEX_MAX_K) return false;
        D.len = len; D.k = k;
        for (int i = 0; i < len; i++) {
            const uint8_t c = s[i];
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return false;
            D.pos[(c >> 1) & 3] |= 1u << i;
        }
        return true;
    };
    for (int i = 0; i < g.n_iter; i++) {
        C.starts[i] = g.starts[i];
        if (C.has_up && !prep(g.up[i], g.up_len[i], g.miss_up, C.up[i])) return false;
        if (C.has_down && !prep(g.down[i], g.down_len[i], g.miss_down, C.down[i])) return false;
    }
    C.max_k = delim ? ((C.has_up ? g.miss_up : 0) > (C.has_down ? g.miss_down : 0) ? (C.has_up ? g.miss_up : 0) : (C.has_down ? g.miss_down : 0)) : 0;
    C.eligible = 1;
    return true;
}

}  // namespace f2q
