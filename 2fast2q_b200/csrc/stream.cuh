// stream.cuh — chunk boundary handling, entirely on the device and stream-ordered (no host round trip):
//   k_prepare  before the tile kernel: completes the partial record carried from the previous chunk with the head
//              of the new chunk ("stitch") and sets the record-aligned start of the range the tile kernel parses
//   k_carry    after the tile kernel: saves the new partial tail record
// Records are "every 4 lines from byte 0" (fast2q.py:324-328), so a chunk cut anywhere needs exactly this much
// state: the bytes of the unfinished record (their newline count mod 4 is the line phase).
#pragma once

#include "f2q_dev.cuh"

namespace f2q {

constexpr int PREP_THREADS = 256;

// chunk = user bytes at buf[delta, delta+n); carry = ctx buffer of carry_cap bytes
__global__ void __launch_bounds__(PREP_THREADS) k_prepare(DevState* S, const uint8_t* __restrict__ buf, uint64_t delta, uint64_t n,
                                                          uint32_t is_last, uint8_t* __restrict__ carry, uint64_t carry_cap,
                                                          uint32_t* tickets, uint32_t q_cap, uint32_t g_cap) {
    __shared__ uint64_t s_h;          // bytes of the chunk head that complete the carried record
    __shared__ uint32_t s_found, s_cnt;
    const uint32_t tid = threadIdx.x;
    const uint32_t tl = S->tail_len, tnl = S->tail_nl;
    __syncthreads();
    if (tid == 0) {
        S->q_count = 0; S->g_count = 0; S->q_cap = q_cap; S->g_cap = g_cap;
        S->last_rec_end = 0; S->nl_total = 0; S->stitch_len = 0; S->stitch_eof = 0; S->appended = 0;
        S->is_last = is_last; S->end = delta + n; S->beg = delta; S->spec_fail = 0; S->spec_ok = 0;
        tickets[0] = 0; tickets[1] = 0;
        s_found = 0; s_cnt = 0; s_h = 0;
    }
    __syncthreads();
    if (tl == 0) return;

    // find the (4 - tnl)-th newline of the chunk, one warp, 32 x 16 bytes per step
    const uint32_t need = 4 - tnl;
    if (tid < 32) {
        uint32_t seen = 0; uint64_t pos = 0; bool found = false; uint64_t h = 0;
        while (pos < n && !found) {
            const uint64_t a = pos + (uint64_t)tid * 16;
            uint32_t mask = 0;
            for (int b = 0; b < 16; b++) if (a + b < n && buf[delta + a + b] == '\n') mask |= 1u << b;
            const uint32_t c = __popc(mask);
            uint32_t incl = c;
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= (uint32_t)d) incl += y; }
            const uint32_t before = seen + incl - c;
            const bool mine = (before < need) && (before + c >= need);
            const uint32_t who = __ballot_sync(0xffffffffu, mine);
            if (who) {
                if (mine) {
                    uint32_t k = need - before;            // the k-th set bit of mask
                    uint32_t m = mask;
                    while (--k) m &= m - 1;
                    h = a + (__ffs(m) - 1) + 1;
                }
                h = __shfl_sync(0xffffffffu, h, __ffs(who) - 1);
                found = true;
            } else {
                seen += __shfl_sync(0xffffffffu, incl, 31);
                pos += 512;
            }
        }
        if (tid == 0) { s_found = found; s_h = h; s_cnt = seen; }
    }
    __syncthreads();
    const bool found = s_found != 0;
    const uint64_t h = found ? s_h : n;                     // bytes of the chunk that go behind the carried tail
    if ((uint64_t)tl + h > carry_cap) {
        if (tid == 0) { atomicOr(&S->error, ERR_RECORD_TOO_LONG); S->tail_len = 0; S->tail_nl = 0; S->beg = S->end; S->appended = 1; }
        return;
    }
    for (uint64_t i = tid; i < h; i += PREP_THREADS) carry[tl + i] = buf[delta + i];
    if (tid == 0) {
        if (found) {
            S->stitch_len = (uint32_t)(tl + h); S->stitch_eof = 0;
            S->beg = delta + h; S->tail_len = 0; S->tail_nl = 0;
        } else {
            // the chunk did not finish the record: keep accumulating; at end of stream parse what we have
            S->tail_len = (uint32_t)(tl + n); S->tail_nl = tnl + s_cnt;
            S->beg = S->end; S->appended = 1;
            if (is_last) { S->stitch_len = (uint32_t)(tl + n); S->stitch_eof = 1; S->tail_len = 0; S->tail_nl = 0; }
        }
    }
}

__global__ void __launch_bounds__(PREP_THREADS) k_carry(DevState* S, const uint8_t* __restrict__ buf, uint8_t* __restrict__ carry,
                                                        uint64_t carry_cap) {
    const uint32_t tid = threadIdx.x;
    if (S->appended) return;
    if (S->is_last) { if (tid == 0) { S->tail_len = 0; S->tail_nl = 0; } return; }
    const uint64_t le = S->last_rec_end;
    const uint64_t tb = le > S->beg ? le : S->beg;
    const uint64_t tlen = S->end - tb;
    const uint32_t nl = S->nl_total & 3u;
    __syncthreads();
    if (tlen > carry_cap) { if (tid == 0) { atomicOr(&S->error, ERR_RECORD_TOO_LONG); S->tail_len = 0; S->tail_nl = 0; } return; }
    for (uint64_t i = tid; i < tlen; i += PREP_THREADS) carry[i] = buf[tb + i];
    if (tid == 0) { S->tail_len = (uint32_t)tlen; S->tail_nl = nl; }
}

// ------------------------------------------------------------------------------------------------
// K0: synthetic FASTQ generator (bench / tests).  Bit-identical to 2fast2q_b200/synth.py:fixed_reads().
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t sm_fin(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_synth(f2q_synth_spec sp, const uint8_t* __restrict__ guides, uint8_t* __restrict__ out) {
    const uint64_t GOLD = 0x9E3779B97F4A7C15ull, K2 = 0xD1342543DE82EF95ull;
    const uint32_t L = sp.read_len, F = sp.feat_len;
    const uint64_t rec = 2ull * L + 18;
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < sp.n_reads; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = sp.first_read + k;
        uint8_t* o = out + k * rec;
        const uint64_t b = sm_fin((sp.seed + 1) * GOLD + i * K2);
        const uint64_t r0 = sm_fin(b + 1 * GOLD), r1 = sm_fin(b + 2 * GOLD);
        const uint32_t cls = (uint32_t)(r0 & 0xFFFF), lowsel = (uint32_t)((r0 >> 16) & 0xFFFF);
        const uint32_t lowpos = (uint32_t)((((r0 >> 32) & 0xFFFF) * L) >> 16);
        const uint32_t lowq = 2 + (uint32_t)((((r0 >> 48) & 0xFFFF) * 27) >> 16);
        const uint32_t gi = (uint32_t)(((r1 & 0xFFFFFFFFull) * sp.n_guides) >> 32);
        o[0] = '@'; o[1] = 'S';
        uint64_t v = i;
        for (int d = 0; d < 11; d++) { o[12 - d] = (uint8_t)('0' + v % 10); v /= 10; }
        o[13] = '\n';
        uint8_t* s = o + 14;
        const char ACGT[4] = {'A', 'C', 'G', 'T'};
        for (uint32_t j = 0; j < F; j++) s[j] = guides[(uint64_t)gi * F + j];
        const uint32_t a = (uint32_t)((r1 >> 32) & 0xFF), bb = (uint32_t)((r1 >> 40) & 0xFF), cc = (uint32_t)((r1 >> 48) & 0xFF);
        const uint32_t sb = (uint32_t)((r1 >> 56) & 0xFF);
        const uint32_t p0 = a % F, d1 = 1 + bb % (F - 1), p1 = (p0 + d1) % F;
        uint32_t d2 = 1 + cc % (F - 2); d2 += (d2 >= d1);
        const uint32_t p2 = (p0 + d2) % F;
        auto code_of = [](uint8_t c) -> uint32_t { return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u; };
        auto subst = [&](uint32_t pos, uint32_t delta) { s[pos] = ACGT[(code_of(s[pos]) + 1 + delta % 3) % 4]; };
        const bool is1 = cls >= sp.cum_exact && cls < sp.cum_sub1, is2 = cls >= sp.cum_sub1 && cls < sp.cum_sub2;
        const bool is3 = cls >= sp.cum_sub2 && cls < sp.cum_sub3, isn = cls >= sp.cum_sub3 && cls < sp.cum_n, isr = cls >= sp.cum_n;
        if (is1 || is2 || is3) subst(p0, sb & 3);
        if (is2 || is3) subst(p1, (sb >> 2) & 3);
        if (is3) subst(p2, (sb >> 4) & 3);
        if (isn) s[(sb * F) >> 8] = 'N';
        if (isr) { const uint64_t rr = sm_fin(b + 5 * GOLD); for (uint32_t j = 0; j < F; j++) s[j] = ACGT[(rr >> (2 * j)) & 3]; }
        const uint64_t t0 = sm_fin(b + 3 * GOLD), t1 = sm_fin(b + 4 * GOLD);
        for (uint32_t j = 0; j < L - F; j++) { const uint64_t src = j < 32 ? t0 : t1; s[F + j] = ACGT[(src >> (2 * (j % 32))) & 3]; }
        s[L] = '\n'; s[L + 1] = '+'; s[L + 2] = '\n';
        uint8_t* q = s + L + 3;
        for (uint32_t w = 0; w < (L + 7) / 8; w++) {
            const uint64_t rq = sm_fin(b + (9 + w) * GOLD);
            for (uint32_t j = 0; j < 8 && w * 8 + j < L; j++) q[w * 8 + j] = (uint8_t)(63 + ((((rq >> (8 * j)) & 0xFF) * 11) >> 8));
        }
        if (lowsel < sp.lowq_per_65536) q[lowpos] = (uint8_t)(33 + lowq);
        q[L] = '\n';
    }
}

}  // namespace f2q
