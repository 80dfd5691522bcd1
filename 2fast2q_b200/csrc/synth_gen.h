// synth_gen.h — K0, the counter-based synthetic FASTQ generator (bench / tests only).
//
// One function writes record i of a workload; the CUDA kernel k_synth (stream.cuh) calls it with one thread per read, and
// tests/native/hostcheck.cpp compiles the SAME function with g++ so that the numpy restatement (2fast2q_b200/synth.py)
// can be checked against it bit for bit without a GPU.  All randomness is splitmix64 of (seed, read index, field).
//
//   shape 0  guide(feat_len) at offset 0 + random tail            configs 2, 3 and the north-star shape (fixed_reads)
//   shape 1  Bar-seq: stagger(0-7) + US + barcode + DS + pad      config 4 (Extract + Count between delimiters)
//   shape 2  dual fixed: X at 0, Y at feat_len + 10               config 5a
//   shape 3  dual delimiters: stagger + U1 X D1 + gap + U2 Y D2   config 5b
// Every record is "@S%011d\n" SEQ "\n+\n" QUAL "\n" = 2L+18 bytes.
#pragma once

#include <stdint.h>

#include "../../include/f2q.h"

#if defined(__CUDACC__)
#define F2Q_HD __host__ __device__ __forceinline__
#else
#define F2Q_HD inline
#endif

namespace f2q {

F2Q_HD uint64_t sm_fin(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// class thresholds (out of 65536) of shapes 1-3
constexpr uint32_t SY1_SUB = 3277, SY1_LACK = 5243, SY1_LEN = 6554;      // 5 % delimiter substitution, 3 % delimiter missing, 2 % 19/21-mers
constexpr uint32_t SY2_XMUT = 6554, SY2_YMUT = 9830, SY2_XRAND = 11796;   // 10 % / 5 % / 3 %
constexpr uint32_t SY2_MISPAIR = 6554;                                    // 10 % of reads pair X with another Y

F2Q_HD uint32_t synth_code_of(uint8_t c) { return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u; }
F2Q_HD uint8_t synth_base(uint32_t code) { return (uint8_t)("ACGT"[code & 3u]); }
F2Q_HD uint8_t synth_subst(uint8_t c, uint32_t delta) { return synth_base((synth_code_of(c) + 1u + delta % 3u) % 4u); }

// writes record i (2L+18 bytes) at o.  guides: shape 0/1 n_guides x feat_len bytes; shapes 2/3: n_guides X's then n_guides Y's
F2Q_HD void synth_record(const f2q_synth_spec& sp, const uint8_t* guides, uint64_t i, uint8_t* o) {
    const uint64_t GOLD = 0x9E3779B97F4A7C15ull, K2 = 0xD1342543DE82EF95ull;
    const uint32_t L = sp.read_len, F = sp.feat_len;
    const uint64_t b = sm_fin((sp.seed + 1) * GOLD + i * K2);
    const uint64_t r0 = sm_fin(b + 1 * GOLD), r1 = sm_fin(b + 2 * GOLD);
    const uint32_t cls = (uint32_t)(r0 & 0xFFFF), lowsel = (uint32_t)((r0 >> 16) & 0xFFFF);
    const uint32_t lowpos = (uint32_t)((((r0 >> 32) & 0xFFFF) * L) >> 16);
    const uint32_t lowq = 2 + (uint32_t)((((r0 >> 48) & 0xFFFF) * 27) >> 16);
    o[0] = '@'; o[1] = 'S';
    uint64_t v = i;
    for (int d = 0; d < 11; d++) { o[12 - d] = (uint8_t)('0' + v % 10); v /= 10; }
    o[13] = '\n';
    uint8_t* s = o + 14;
    if (sp.shape == 0) {
        const uint32_t gi = (uint32_t)(((r1 & 0xFFFFFFFFull) * sp.n_guides) >> 32);
        for (uint32_t j = 0; j < F; j++) s[j] = guides[(uint64_t)gi * F + j];
        const uint32_t a = (uint32_t)((r1 >> 32) & 0xFF), bb = (uint32_t)((r1 >> 40) & 0xFF), cc = (uint32_t)((r1 >> 48) & 0xFF);
        const uint32_t sb = (uint32_t)((r1 >> 56) & 0xFF);
        const uint32_t p0 = a % F, d1 = 1 + bb % (F - 1), p1 = (p0 + d1) % F;
        uint32_t d2 = 1 + cc % (F - 2); d2 += (d2 >= d1);
        const uint32_t p2 = (p0 + d2) % F;
        const bool is1 = cls >= sp.cum_exact && cls < sp.cum_sub1, is2 = cls >= sp.cum_sub1 && cls < sp.cum_sub2;
        const bool is3 = cls >= sp.cum_sub2 && cls < sp.cum_sub3, isn = cls >= sp.cum_sub3 && cls < sp.cum_n, isr = cls >= sp.cum_n;
        if (is1 || is2 || is3) s[p0] = synth_subst(s[p0], sb & 3);
        if (is2 || is3) s[p1] = synth_subst(s[p1], (sb >> 2) & 3);
        if (is3) s[p2] = synth_subst(s[p2], (sb >> 4) & 3);
        if (isn) s[(sb * F) >> 8] = 'N';
        if (isr) { const uint64_t rr = sm_fin(b + 5 * GOLD); for (uint32_t j = 0; j < F; j++) s[j] = synth_base((uint32_t)(rr >> (2 * j))); }
        const uint64_t t0 = sm_fin(b + 3 * GOLD), t1 = sm_fin(b + 4 * GOLD);
        for (uint32_t j = 0; j < L - F; j++) { const uint64_t src = j < 32 ? t0 : t1; s[F + j] = synth_base((uint32_t)(src >> (2 * (j % 32)))); }
    } else {
        // background: L random bases; the segments below overwrite parts of it
        for (uint32_t w = 0; w * 32 < L; w++) {
            const uint64_t bg = sm_fin(b + (uint64_t)(40 + w) * GOLD);
            for (uint32_t j = 0; j < 32 && w * 32 + j < L; j++) s[w * 32 + j] = synth_base((uint32_t)(bg >> (2 * j)));
        }
        const uint64_t r2 = sm_fin(b + 6 * GOLD);
        uint32_t p = 0;
        // copy n bytes of src to s[p ..), optionally with one substitution at index mpos (mpos >= n: none); clipped at L
        auto put = [&](const uint8_t* src, uint32_t n, uint32_t mpos, uint32_t mdelta) {
            for (uint32_t k = 0; k < n; k++, p++) {
                if (p >= L) continue;
                uint8_t c = src[k];
                if (k == mpos) c = synth_subst(c, mdelta);
                s[p] = c;
            }
        };
        const uint32_t NONE = 0xFFFFFFFFu;
        if (sp.shape == 1) {
            const uint64_t u32 = r1 & 0xFFFFFFFFull;
            const uint64_t sq = (u32 * u32) >> 32;                     // u^2: heavy-tailed abundance, low indices are frequent
            const uint32_t bi = (uint32_t)((sq * sp.n_guides) >> 32);
            const uint32_t which = (uint32_t)((r2 >> 3) & 1), mdelta = (uint32_t)((r2 >> 16) & 3);
            const bool sub = cls < SY1_SUB, lack = cls >= SY1_SUB && cls < SY1_LACK, odd = cls >= SY1_LACK && cls < SY1_LEN;
            const uint32_t ul = sp.delim_len[0], dl = sp.delim_len[1];
            p = (uint32_t)(r2 & 7);
            if (lack && which == 0) p += ul;
            else put(sp.delim[0], ul, (sub && which == 0) ? (uint32_t)((((r2 >> 8) & 0xFF) * ul) >> 8) : NONE, mdelta);
            uint32_t bl = F;
            if (odd && ((r2 >> 4) & 1)) bl = F - 1;
            put(guides + (uint64_t)bi * F, bl, NONE, 0);
            if (odd && !((r2 >> 4) & 1)) p += 1;                       // a 21-mer: one background base behind the barcode
            if (lack && which == 1) p += dl;
            else put(sp.delim[1], dl, (sub && which == 1) ? (uint32_t)((((r2 >> 8) & 0xFF) * dl) >> 8) : NONE, mdelta);
        } else {
            const uint32_t n = sp.n_guides;
            const uint32_t k = (uint32_t)(((r1 & 0xFFFFFFFFull) * n) >> 32);
            const uint32_t ky = ((uint32_t)(r2 & 0xFFFF) < SY2_MISPAIR) ? (uint32_t)(((r2 >> 32) * n) >> 32) : k;
            const uint8_t* X = guides + (uint64_t)k * F;
            const uint8_t* Y = guides + ((uint64_t)n + ky) * F;
            const uint32_t mpos = (uint32_t)((((r2 >> 16) & 0xFF) * F) >> 8), mdelta = (uint32_t)((r2 >> 24) & 3);
            const bool xmut = cls < SY2_XMUT, ymut = cls >= SY2_XMUT && cls < SY2_YMUT, xrand = cls >= SY2_YMUT && cls < SY2_XRAND;
            if (sp.shape == 2) {
                if (xrand) p += F; else put(X, F, xmut ? mpos : NONE, mdelta);
                p += 10;
                put(Y, F, ymut ? mpos : NONE, mdelta);
            } else {
                p = (uint32_t)((r2 >> 26) & 3);
                put(sp.delim[0], sp.delim_len[0], NONE, 0);
                if (xrand) p += F; else put(X, F, xmut ? mpos : NONE, mdelta);
                put(sp.delim[1], sp.delim_len[1], NONE, 0);
                p += (uint32_t)((r2 >> 28) & 3) % 3u;
                put(sp.delim[2], sp.delim_len[2], NONE, 0);
                put(Y, F, ymut ? mpos : NONE, mdelta);
                put(sp.delim[3], sp.delim_len[3], NONE, 0);
            }
        }
    }
    s[L] = '\n'; s[L + 1] = '+'; s[L + 2] = '\n';
    uint8_t* q = s + L + 3;
    for (uint32_t w = 0; w < (L + 7) / 8; w++) {
        const uint64_t rq = sm_fin(b + (9 + w) * GOLD);
        for (uint32_t j = 0; j < 8 && w * 8 + j < L; j++) q[w * 8 + j] = (uint8_t)(63 + ((((rq >> (8 * j)) & 0xFF) * 11) >> 8));
    }
    if (lowsel < sp.lowq_per_65536) q[lowpos] = (uint8_t)(33 + lowq);
    q[L] = '\n';
}

}  // namespace f2q
