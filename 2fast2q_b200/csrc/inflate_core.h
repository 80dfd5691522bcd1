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

}  // namespace f2q
