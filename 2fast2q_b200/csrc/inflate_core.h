// inflate_core.h — DEFLATE (RFC 1951) decoder for ONE raw deflate stream per thread: the payload of a BGZF block
// (bgzip writes gzip members of <= 64 KiB that carry their compressed size in an extra field and their uncompressed size in
// their trailer, so the blocks of a file are independent and their output offsets are known before anything is inflated).
//
// SURVEY.md §8(f)-1 / VERDICT r1 task 9: the reference inflates with `gzip.open` on one host thread (fast2q.py:568); here a
// chunk of COMPRESSED bytes crosses PCIe (3-4x fewer bytes than the FASTQ text) and k_inflate_bgzf (stream.cuh) decodes
// thousands of blocks at once, one thread each, straight into the chunk buffer the streaming kernel parses.
//
// The decoder is the canonical-Huffman, bit-serial one (code lengths -> counts + sorted symbols; a symbol is decoded by walking
// the code lengths): small per-thread state (about 1 KB of local memory), no shared tables, and FASTQ text has short codes.
// It is a pure function compiled for the device by nvcc and for the host by g++: tests/native/hostcheck.cpp checks it
// against zlib on the CPU (stored, fixed and dynamic blocks, every compression level), the -m gpu tests against the same
// files through the kernel.
#pragma once

#include <stdint.h>

#include "synth_gen.h"      // F2Q_HD

namespace f2q {

constexpr int INFL_MAXBITS = 15, INFL_MAXLCODES = 286, INFL_MAXDCODES = 30, INFL_FIXLCODES = 288;

struct InflState {
    const uint8_t* in; uint32_t in_len, in_pos;
    uint64_t bitbuf; uint32_t bitcnt;
    uint8_t* out; uint32_t out_len, out_pos;
    int err;                 // sticky: 1 input exhausted, 2 bad data, 3 output overflow
};

struct InflHuff {
    uint16_t count[INFL_MAXBITS + 1];
    uint16_t symbol[INFL_FIXLCODES];
};
struct InflHuffD {
    uint16_t count[INFL_MAXBITS + 1];
    uint16_t symbol[INFL_MAXDCODES];
};

F2Q_HD void infl_refill(InflState& s) {
    while (s.bitcnt <= 56u && s.in_pos < s.in_len) { s.bitbuf |= (uint64_t)s.in[s.in_pos++] << s.bitcnt; s.bitcnt += 8u; }
}
// n <= 16 bits, LSB first
F2Q_HD uint32_t infl_bits(InflState& s, uint32_t n) {
    if (s.bitcnt < n) { infl_refill(s); if (s.bitcnt < n) { s.err = s.err ? s.err : 1; return 0; } }
    const uint32_t v = (uint32_t)(s.bitbuf & ((1ull << n) - 1ull));
    s.bitbuf >>= n; s.bitcnt -= n;
    return v;
}

// one symbol of a canonical code (count[len] codes of each length, symbols sorted by code)
template <class H>
F2Q_HD int infl_decode(InflState& s, const H& h) {
    if (s.bitcnt < (uint32_t)INFL_MAXBITS) infl_refill(s);
    int code = 0, first = 0, index = 0;
    uint64_t buf = s.bitbuf;
    const uint32_t have = s.bitcnt;
    for (uint32_t len = 1; len <= (uint32_t)INFL_MAXBITS; len++) {
        if (len > have) { s.err = s.err ? s.err : 1; return -1; }
        code |= (int)(buf & 1u); buf >>= 1;
        const int count = h.count[len];
        if (code - count < first) { s.bitbuf = buf; s.bitcnt = have - len; return h.symbol[index + (code - first)]; }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    s.err = s.err ? s.err : 2;
    return -1;
}

// code lengths -> canonical decoding tables; returns 0 complete code, < 0 over-subscribed, > 0 incomplete
template <class H>
F2Q_HD int infl_construct(H& h, const uint8_t* length, int n) {
    for (int len = 0; len <= INFL_MAXBITS; len++) h.count[len] = 0;
    for (int sym = 0; sym < n; sym++) h.count[length[sym]]++;
    if (h.count[0] == n) return 0;
    int left = 1;
    for (int len = 1; len <= INFL_MAXBITS; len++) { left <<= 1; left -= h.count[len]; if (left < 0) return left; }
    uint16_t offs[INFL_MAXBITS + 1];
    offs[1] = 0;
    for (int len = 1; len < INFL_MAXBITS; len++) offs[len + 1] = (uint16_t)(offs[len] + h.count[len]);
    for (int sym = 0; sym < n; sym++) if (length[sym] != 0) h.symbol[offs[length[sym]]++] = (uint16_t)sym;
    return left;
}

F2Q_HD void infl_codes(InflState& s, const InflHuff& lencode, const InflHuffD& distcode) {
    static const uint16_t lens[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dists[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (;;) {
        int symbol = infl_decode(s, lencode);
        if (symbol < 0) return;
        if (symbol < 256) {
            if (s.out_pos >= s.out_len) { s.err = s.err ? s.err : 3; return; }
            s.out[s.out_pos++] = (uint8_t)symbol;
        } else if (symbol == 256) return;
        else {
            symbol -= 257;
            if (symbol >= 29) { s.err = s.err ? s.err : 2; return; }
            const uint32_t len = lens[symbol] + infl_bits(s, lext[symbol]);
            const int ds = infl_decode(s, distcode);
            if (ds < 0) return;
            if (ds >= 30) { s.err = s.err ? s.err : 2; return; }
            const uint32_t dist = dists[ds] + infl_bits(s, dext[ds]);
            if (s.err) return;
            if (dist > s.out_pos) { s.err = 2; return; }
            if (s.out_pos + len > s.out_len) { s.err = 3; return; }
            uint8_t* o = s.out + s.out_pos;
            const uint8_t* f = o - dist;
            for (uint32_t k = 0; k < len; k++) o[k] = f[k];
            s.out_pos += len;
        }
    }
}

// inflates one raw deflate stream; returns 0 when exactly out_len bytes were produced and the final block ended, else the error
F2Q_HD int inflate_raw(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len) {
    InflState s;
    s.in = in; s.in_len = in_len; s.in_pos = 0; s.bitbuf = 0; s.bitcnt = 0; s.out = out; s.out_len = out_len; s.out_pos = 0; s.err = 0;
    InflHuff lencode;
    InflHuffD distcode;
    uint8_t lengths[INFL_MAXLCODES + INFL_MAXDCODES + 4];
    int last;
    do {
        last = (int)infl_bits(s, 1);
        const uint32_t type = infl_bits(s, 2);
        if (s.err) break;
        if (type == 0) {
            // stored: skip to the byte boundary, LEN / NLEN, raw bytes
            const uint32_t drop = s.bitcnt & 7u;
            s.bitbuf >>= drop; s.bitcnt -= drop;
            const uint32_t len = infl_bits(s, 16), nlen = infl_bits(s, 16);
            if (s.err) break;
            if ((len ^ 0xFFFFu) != nlen) { s.err = 2; break; }
            if (s.out_pos + len > s.out_len) { s.err = 3; break; }
            for (uint32_t k = 0; k < len; k++) {
                const uint32_t b = infl_bits(s, 8);
                if (s.err) break;
                s.out[s.out_pos++] = (uint8_t)b;
            }
        } else if (type == 1) {
            int sym = 0;
            for (; sym < 144; sym++) lengths[sym] = 8;
            for (; sym < 256; sym++) lengths[sym] = 9;
            for (; sym < 280; sym++) lengths[sym] = 7;
            for (; sym < INFL_FIXLCODES; sym++) lengths[sym] = 8;
            infl_construct(lencode, lengths, INFL_FIXLCODES);
            for (sym = 0; sym < INFL_MAXDCODES; sym++) lengths[sym] = 5;
            infl_construct(distcode, lengths, INFL_MAXDCODES);
            infl_codes(s, lencode, distcode);
        } else if (type == 2) {
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            const int nlen = (int)infl_bits(s, 5) + 257, ndist = (int)infl_bits(s, 5) + 1, ncode = (int)infl_bits(s, 4) + 4;
            if (s.err) break;
            if (nlen > INFL_MAXLCODES || ndist > INFL_MAXDCODES) { s.err = 2; break; }
            int index = 0;
            for (; index < ncode; index++) lengths[order[index]] = (uint8_t)infl_bits(s, 3);
            for (; index < 19; index++) lengths[order[index]] = 0;
            if (s.err) break;
            if (infl_construct(lencode, lengths, 19) != 0) { s.err = 2; break; }
            index = 0;
            while (index < nlen + ndist) {
                int symbol = infl_decode(s, lencode);
                if (symbol < 0) break;
                if (symbol < 16) lengths[index++] = (uint8_t)symbol;
                else {
                    int len = 0, rep;
                    if (symbol == 16) { if (index == 0) { s.err = 2; break; } len = lengths[index - 1]; rep = 3 + (int)infl_bits(s, 2); }
                    else if (symbol == 17) rep = 3 + (int)infl_bits(s, 3);
                    else rep = 11 + (int)infl_bits(s, 7);
                    if (s.err) break;
                    if (index + rep > nlen + ndist) { s.err = 2; break; }
                    while (rep--) lengths[index++] = (uint8_t)len;
                }
            }
            if (s.err) break;
            if (lengths[256] == 0) { s.err = 2; break; }
            int e = infl_construct(lencode, lengths, nlen);
            if (e < 0 || (e > 0 && nlen - lencode.count[0] != 1)) { s.err = 2; break; }
            e = infl_construct(distcode, lengths + nlen, ndist);
            if (e < 0 || (e > 0 && ndist - distcode.count[0] != 1)) { s.err = 2; break; }
            infl_codes(s, lencode, distcode);
        } else s.err = 2;
    } while (!last && !s.err);
    if (!s.err && s.out_pos != s.out_len) s.err = 2;
    return s.err;
}


// ------------------------------------------------------------------------------------------------------------
// The same decoder as a LOCK-STEP state machine, for a warp whose 32 lanes inflate 32 blocks.
// inflate_raw above lets every lane follow its own control flow; measured on the device (ncu: 1.01 active threads per
// warp instruction) the lanes of a warp then run one after the other.  Here every lane executes the SAME loop body; what a
// lane does in an iteration depends on its state: read a block header and build its tables (rare, the only part in which
// lanes wait for each other), decode one literal / length+distance pair through a lookup table, copy up to 8 bytes of a
// match, copy up to 8 bytes of a stored block.  Codes longer than the table's index are decoded bit-serially from the
// canonical tables (FASTQ text has none).
// lut / dlut: the lane's lookup tables, element i of lane L at lut[i * stride + L] (shared memory on the device: the
// interleaving makes the lanes' random look-ups hit distinct banks); entry = symbol << 4 | code length, 0 = no short code.
// ------------------------------------------------------------------------------------------------------------
constexpr int INFL_LUT_BITS = 9, INFL_DLUT_BITS = 7;
enum { INFL_ST_HEADER = 0, INFL_ST_SYMBOL = 1, INFL_ST_COPY = 2, INFL_ST_STORED = 3, INFL_ST_DONE = 4 };

struct InflLane {
    InflState s;
    InflHuff lencode;
    InflHuffD distcode;
    uint32_t state, last, copy_len, copy_dist;
};

F2Q_HD uint32_t infl_rev(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// lookup table of a canonical code: every index whose low `len` bits are the (bit-reversed) code of a symbol of length
// len <= bits gets symbol << 4 | len
template <class H>
F2Q_HD void infl_build_lut(const H& h, int nsym_max, uint16_t* lut, uint32_t stride, int bits) {
    for (uint32_t i = 0; i < (1u << bits); i++) lut[i * stride] = 0;
    uint32_t code = 0, index = 0;
    for (int len = 1; len <= INFL_MAXBITS; len++) {
        const uint32_t cnt = h.count[len];
        if (len <= bits) {
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t sym = h.symbol[index + k], r = infl_rev(code + k, len);
                for (uint32_t hi = 0; hi < (1u << (bits - len)); hi++) lut[(r | (hi << len)) * stride] = (uint16_t)((sym << 4) | (uint32_t)len);
            }
        }
        code = (code + cnt) << 1; index += cnt;
    }
    (void)nsym_max;
}

// 4 more input bytes when there is room for them (bytes behind the stream's end read as what follows it in memory: a valid
// stream never uses them, and the final check refuses a stream that did)
F2Q_HD void infl_refill4(InflState& s) {
    if (s.bitcnt <= 32u) {
        const uint8_t* p = s.in + s.in_pos;
        const uint32_t w = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        s.bitbuf |= (uint64_t)w << s.bitcnt; s.bitcnt += 32u; s.in_pos += 4u;
    }
}
F2Q_HD uint32_t infl_take(InflState& s, uint32_t n) {
    const uint32_t v = (uint32_t)(s.bitbuf & ((1ull << n) - 1ull));
    s.bitbuf >>= n; s.bitcnt -= n;
    return v;
}
template <class H>
F2Q_HD int infl_decode_lut(InflState& s, const H& h, const uint16_t* lut, uint32_t stride, int bits) {
    const uint32_t e = lut[((uint32_t)s.bitbuf & ((1u << bits) - 1u)) * stride];
    if (e & 15u) { s.bitbuf >>= (e & 15u); s.bitcnt -= (e & 15u); return (int)(e >> 4); }
    // a code longer than the table's index: bit-serial from the canonical tables (at least 33 bits are buffered)
    int code = 0, first = 0, index = 0;
    uint64_t buf = s.bitbuf;
    for (uint32_t len = 1; len <= (uint32_t)INFL_MAXBITS; len++) {
        code |= (int)(buf & 1u); buf >>= 1;
        const int count = h.count[len];
        if (code - count < first) { s.bitbuf = buf; s.bitcnt -= len; return h.symbol[index + (code - first)]; }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    s.err = s.err ? s.err : 2;
    return -1;
}

// block header of one lane (the divergent part): stored -> INFL_ST_STORED; fixed / dynamic -> tables + INFL_ST_SYMBOL
F2Q_HD void infl_header(InflLane& L, uint16_t* lut, uint16_t* dlut, uint32_t stride) {
    InflState& s = L.s;
    uint8_t lengths[INFL_MAXLCODES + INFL_MAXDCODES + 4];
    infl_refill4(s);
    L.last = infl_take(s, 1);
    const uint32_t type = infl_take(s, 2);
    if (type == 0) {
        const uint32_t drop = s.bitcnt & 7u;
        s.bitbuf >>= drop; s.bitcnt -= drop;
        infl_refill4(s);
        const uint32_t len = infl_take(s, 16), nlen = infl_take(s, 16);
        if ((len ^ 0xFFFFu) != nlen) { s.err = 2; return; }
        if (s.out_pos + len > s.out_len) { s.err = 3; return; }
        L.copy_len = len; L.state = INFL_ST_STORED;
        return;
    }
    if (type == 1) {
        int sym = 0;
        for (; sym < 144; sym++) lengths[sym] = 8;
        for (; sym < 256; sym++) lengths[sym] = 9;
        for (; sym < 280; sym++) lengths[sym] = 7;
        for (; sym < INFL_FIXLCODES; sym++) lengths[sym] = 8;
        infl_construct(L.lencode, lengths, INFL_FIXLCODES);
        for (sym = 0; sym < INFL_MAXDCODES; sym++) lengths[sym] = 5;
        infl_construct(L.distcode, lengths, INFL_MAXDCODES);
    } else if (type == 2) {
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        infl_refill4(s);
        const int nlen = (int)infl_take(s, 5) + 257, ndist = (int)infl_take(s, 5) + 1, ncode = (int)infl_take(s, 4) + 4;
        if (nlen > INFL_MAXLCODES || ndist > INFL_MAXDCODES) { s.err = 2; return; }
        int index = 0;
        for (; index < ncode; index++) { infl_refill4(s); lengths[order[index]] = (uint8_t)infl_take(s, 3); }
        for (; index < 19; index++) lengths[order[index]] = 0;
        if (infl_construct(L.lencode, lengths, 19) != 0) { s.err = 2; return; }
        index = 0;
        while (index < nlen + ndist) {
            infl_refill4(s);
            // (the code-length code has at most 19 symbols of <= 7 bits: the bit-serial decoder, on buffered bits)
            int code = 0, first = 0, idx2 = 0, symbol = -1;
            for (uint32_t len = 1; len <= 7u; len++) {
                code |= (int)(s.bitbuf & 1u); s.bitbuf >>= 1; s.bitcnt -= 1;
                const int count = L.lencode.count[len];
                if (code - count < first) { symbol = L.lencode.symbol[idx2 + (code - first)]; break; }
                idx2 += count; first += count; first <<= 1; code <<= 1;
            }
            if (symbol < 0) { s.err = 2; return; }
            if (symbol < 16) lengths[index++] = (uint8_t)symbol;
            else {
                int len = 0, rep;
                if (symbol == 16) { if (index == 0) { s.err = 2; return; } len = lengths[index - 1]; rep = 3 + (int)infl_take(s, 2); }
                else if (symbol == 17) rep = 3 + (int)infl_take(s, 3);
                else rep = 11 + (int)infl_take(s, 7);
                if (index + rep > nlen + ndist) { s.err = 2; return; }
                while (rep--) lengths[index++] = (uint8_t)len;
            }
        }
        if (lengths[256] == 0) { s.err = 2; return; }
        int e = infl_construct(L.lencode, lengths, nlen);
        if (e < 0 || (e > 0 && nlen - L.lencode.count[0] != 1)) { s.err = 2; return; }
        e = infl_construct(L.distcode, lengths + nlen, ndist);
        if (e < 0 || (e > 0 && ndist - L.distcode.count[0] != 1)) { s.err = 2; return; }
    } else { s.err = 2; return; }
    infl_build_lut(L.lencode, INFL_FIXLCODES, lut, stride, INFL_LUT_BITS);
    infl_build_lut(L.distcode, INFL_MAXDCODES, dlut, stride, INFL_DLUT_BITS);
    L.state = INFL_ST_SYMBOL;
}

F2Q_HD void infl_lane_init(InflLane& L, const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len) {
    L.s.in = in; L.s.in_len = in_len; L.s.in_pos = 0; L.s.bitbuf = 0; L.s.bitcnt = 0; L.s.out = out; L.s.out_len = out_len; L.s.out_pos = 0; L.s.err = 0;
    L.state = in_len ? INFL_ST_HEADER : INFL_ST_DONE; L.last = 0; L.copy_len = 0; L.copy_dist = 0;
    if (!in_len && out_len) L.s.err = 2;
}

// one iteration of the state machine for one lane; the caller loops while any lane is not INFL_ST_DONE
F2Q_HD void infl_step(InflLane& L, uint16_t* lut, uint16_t* dlut, uint32_t stride) {
    static const uint16_t lens[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dists[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    InflState& s = L.s;
    if (s.err) { L.state = INFL_ST_DONE; return; }
    if (L.state == INFL_ST_HEADER) { infl_header(L, lut, dlut, stride); if (s.err) L.state = INFL_ST_DONE; return; }
    if (L.state == INFL_ST_SYMBOL) {
        infl_refill4(s);
        int symbol = infl_decode_lut(s, L.lencode, lut, stride, INFL_LUT_BITS);
        if (symbol < 0) { L.state = INFL_ST_DONE; return; }
        if (symbol < 256) {
            if (s.out_pos >= s.out_len) { s.err = 3; L.state = INFL_ST_DONE; return; }
            s.out[s.out_pos++] = (uint8_t)symbol;
        } else if (symbol == 256) L.state = L.last ? INFL_ST_DONE : INFL_ST_HEADER;
        else {
            symbol -= 257;
            if (symbol >= 29) { s.err = 2; L.state = INFL_ST_DONE; return; }
            const uint32_t len = lens[symbol] + infl_take(s, lext[symbol]);
            infl_refill4(s);
            const int ds = infl_decode_lut(s, L.distcode, dlut, stride, INFL_DLUT_BITS);
            if (ds < 0 || ds >= 30) { s.err = 2; L.state = INFL_ST_DONE; return; }
            const uint32_t dist = dists[ds] + infl_take(s, dext[ds]);
            if (dist > s.out_pos || s.out_pos + len > s.out_len) { s.err = dist > s.out_pos ? 2 : 3; L.state = INFL_ST_DONE; return; }
            L.copy_len = len; L.copy_dist = dist; L.state = INFL_ST_COPY;
        }
        return;
    }
    if (L.state == INFL_ST_COPY) {
        const uint32_t n = L.copy_len < 8u ? L.copy_len : 8u;
        uint8_t* o = s.out + s.out_pos;
        const uint8_t* f = o - L.copy_dist;
        for (uint32_t k = 0; k < n; k++) o[k] = f[k];
        s.out_pos += n; L.copy_len -= n;
        if (!L.copy_len) L.state = INFL_ST_SYMBOL;
        return;
    }
    if (L.state == INFL_ST_STORED) {
        // the bit buffer holds whole bytes here
        uint32_t n = L.copy_len < 4u ? L.copy_len : 4u;
        infl_refill4(s);
        for (uint32_t k = 0; k < n; k++) s.out[s.out_pos++] = (uint8_t)infl_take(s, 8);
        L.copy_len -= n;
        if (!L.copy_len) L.state = L.last ? INFL_ST_DONE : INFL_ST_HEADER;
        return;
    }
}

// end of a lane: everything produced, and not a bit taken from behind the stream's end
F2Q_HD int infl_lane_result(const InflLane& L) {
    if (L.s.err) return L.s.err;
    if (L.s.out_pos != L.s.out_len) return 2;
    if ((uint64_t)L.s.in_pos * 8u - L.s.bitcnt > (uint64_t)L.s.in_len * 8u) return 1;
    return 0;
}

}  // namespace f2q
