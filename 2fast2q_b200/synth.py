"""
BENCH/TEST TOOLING (not an oracle, not on the product path) — deterministic synthetic FASTQ generators.

`fixed_reads()` is the numpy restatement of the CUDA generator K0 (`f2q_synth_fastq`,
k_synth in 2fast2q_b200/csrc/stream.cuh): both must emit identical bytes for the same spec, which
tests/test_gpu_parity.py::test_synth_generator_matches_numpy checks.  All randomness comes from a counter-based splitmix64 construction, so
nothing depends on numpy's RNG streams:

    fin(z)    : z ^= z>>30; z *= 0xBF58476D1CE4E5B9; z ^= z>>27; z *= 0x94D049BB133111EB; z ^= z>>31
    base(i)   = fin((seed+1)*GOLD + i*0xD1342543DE82EF95)
    r(i, f)   = fin(base(i) + (f+1)*GOLD)                GOLD = 0x9E3779B97F4A7C15

Record layout (B_read = 2L+18 bytes):  "@S%011d\n" SEQ "\n+\n" QUAL "\n".
"""
from __future__ import annotations

import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
K2 = np.uint64(0xD1342543DE82EF95)
U64 = np.uint64
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

# field indices of r(i, f)
F_MISC, F_GUIDE, F_TAIL0, F_TAIL1, F_RAND, F_QUAL0 = 0, 1, 2, 3, 4, 8


def fin(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> U64(30))) * M1
        z = (z ^ (z >> U64(27))) * M2
        z = z ^ (z >> U64(31))
    return z


def base(seed, i):
    with np.errstate(over="ignore"):
        return fin((U64(seed) + U64(1)) * GOLD + np.asarray(i, dtype=np.uint64) * K2)


def rnd(b, f):
    with np.errstate(over="ignore"):
        return fin(b + U64(f + 1) * GOLD)


def make_library(seed: int, n: int, length: int = 20):
    """n distinct uniform-random `length`-mers (bytes), names sg%06d.  Deterministic in (seed, n, length)."""
    out, seen = [], set()
    i = 0
    while len(out) < n:
        m = max(1024, (n - len(out)) * 2)
        idx = np.arange(i, i + m, dtype=np.uint64)
        i += m
        b = base(seed ^ 0x5EED, idx)
        nwords = (length + 31) // 32
        codes = np.zeros((m, nwords * 32), dtype=np.uint8)
        for w in range(nwords):
            r = rnd(b, 100 + w)
            for j in range(32):
                codes[:, w * 32 + j] = ((r >> U64(2 * j)) & U64(3)).astype(np.uint8)
        seqs = ACGT[codes[:, :length]]
        for row in seqs:
            s = row.tobytes()
            if s not in seen:
                seen.add(s)
                out.append(s)
                if len(out) == n:
                    break
    names = ["sg%06d" % k for k in range(n)]
    return names, out


def random_kmers(seed: int, n: int, length: int = 20):
    """n uniform-random `length`-mers (bytes), NOT de-duplicated — the barcode pool of the Bar-seq workload"""
    idx = np.arange(n, dtype=np.uint64)
    b = base(seed ^ 0xBA5C, idx)
    codes = np.zeros((n, ((length + 31) // 32) * 32), dtype=np.uint8)
    for w in range((length + 31) // 32):
        r = rnd(b, 200 + w)
        for j in range(32):
            codes[:, w * 32 + j] = ((r >> U64(2 * j)) & U64(3)).astype(np.uint8)
    seqs = ACGT[codes[:, :length]]
    return [row.tobytes() for row in seqs]


def default_spec(config: int):
    """class mixes of BASELINE.json configs 2 and 3, cumulative thresholds out of 65536"""
    if config == 2:
        fr = dict(exact=0.82, sub1=0.10, sub2=0.03, sub3=0.0, n=0.02)      # rest (3 %) random
        L, seed = 50, 2
    elif config == 3:
        fr = dict(exact=0.75, sub1=0.12, sub2=0.06, sub3=0.02, n=0.01)     # rest (4 %) random
        L, seed = 75, 3
    else:
        raise ValueError(config)
    cum, acc = {}, 0.0
    for k in ("exact", "sub1", "sub2", "sub3", "n"):
        acc += fr[k]
        cum[k] = int(round(acc * 65536))
    return dict(seed=seed, read_len=L, feat_len=20, cum_exact=cum["exact"], cum_sub1=cum["sub1"],
                cum_sub2=cum["sub2"], cum_sub3=cum["sub3"], cum_n=cum["n"], lowq_per_65536=int(round(0.08 * 65536)))


def fixed_reads(guides, first_read: int, n_reads: int, *, seed, read_len, feat_len, cum_exact, cum_sub1, cum_sub2,
                cum_sub3, cum_n, lowq_per_65536):
    """reads [first_read, first_read+n_reads) of the fixed-position workload -> uint8[n_reads*(2L+18)]"""
    L, F = int(read_len), int(feat_len)
    assert 4 <= F <= 32 and F <= L <= 150
    G = np.frombuffer(b"".join(guides), dtype=np.uint8).reshape(len(guides), F)
    code_of = np.zeros(256, dtype=np.uint8)
    for c, ch in enumerate(b"ACGT"):
        code_of[ch] = c
    n = int(n_reads)
    idx = np.arange(first_read, first_read + n, dtype=np.uint64)
    b = base(seed, idx)
    r0, r1 = rnd(b, F_MISC), rnd(b, F_GUIDE)
    cls = (r0 & U64(0xFFFF)).astype(np.int64)
    lowsel = ((r0 >> U64(16)) & U64(0xFFFF)).astype(np.int64)
    lowpos = ((((r0 >> U64(32)) & U64(0xFFFF)) * U64(L)) >> U64(16)).astype(np.int64)
    lowq = (U64(2) + ((((r0 >> U64(48)) & U64(0xFFFF)) * U64(27)) >> U64(16))).astype(np.uint8)
    gi = (((r1 & U64(0xFFFFFFFF)) * U64(len(guides))) >> U64(32)).astype(np.int64)

    rec = 2 * L + 18
    out = np.empty((n, rec), dtype=np.uint8)
    # header "@S%011d\n"
    out[:, 0] = ord("@")
    out[:, 1] = ord("S")
    v = idx.copy()
    for d in range(11):
        out[:, 12 - d] = (v % U64(10)).astype(np.uint8) + ord("0")
        v //= U64(10)
    out[:, 13] = 10
    seq = out[:, 14:14 + L]
    # guide part
    seq[:, :F] = G[gi]
    # substitution slots
    a = ((r1 >> U64(32)) & U64(0xFF)).astype(np.int64)
    bb = ((r1 >> U64(40)) & U64(0xFF)).astype(np.int64)
    cc = ((r1 >> U64(48)) & U64(0xFF)).astype(np.int64)
    sb = ((r1 >> U64(56)) & U64(0xFF)).astype(np.int64)      # substitution deltas (2 bits each) / N position
    p0 = a % F
    d1 = 1 + bb % (F - 1)
    p1 = (p0 + d1) % F
    d2 = 1 + cc % (F - 2)
    d2 = d2 + (d2 >= d1)
    p2 = (p0 + d2) % F
    rows = np.arange(n)
    is1 = (cls >= cum_exact) & (cls < cum_sub1)
    is2 = (cls >= cum_sub1) & (cls < cum_sub2)
    is3 = (cls >= cum_sub2) & (cls < cum_sub3)
    isn = (cls >= cum_sub3) & (cls < cum_n)
    isr = cls >= cum_n

    def subst(mask, pos, delta):
        rr = rows[mask]
        old = code_of[seq[rr, pos[mask]]]
        seq[rr, pos[mask]] = ACGT[(old + 1 + delta[mask] % 3) % 4]

    any1 = is1 | is2 | is3
    subst(any1, p0, sb & 3)
    subst(is2 | is3, p1, (sb >> 2) & 3)
    subst(is3, p2, (sb >> 4) & 3)
    pn = (sb * F) >> 8
    seq[rows[isn], pn[isn]] = ord("N")
    # random 20-mer replaces the guide for the "random" class
    rr4 = rnd(b, F_RAND)
    for j in range(F):
        col = ACGT[((rr4 >> U64(2 * j)) & U64(3)).astype(np.int64)]
        seq[isr, j] = col[isr]
    # random tail
    t0, t1 = rnd(b, F_TAIL0), rnd(b, F_TAIL1)
    for j in range(L - F):
        src = t0 if j < 32 else t1
        seq[:, F + j] = ACGT[((src >> U64(2 * (j % 32))) & U64(3)).astype(np.int64)]
    out[:, 14 + L] = 10
    out[:, 15 + L] = ord("+")
    out[:, 16 + L] = 10
    qual = out[:, 17 + L:17 + 2 * L]
    for w in range((L + 7) // 8):
        rq = rnd(b, F_QUAL0 + w)
        for j in range(8):
            k = w * 8 + j
            if k >= L:
                break
            byte = (rq >> U64(8 * j)) & U64(0xFF)
            qual[:, k] = (U64(63) + ((byte * U64(11)) >> U64(8))).astype(np.uint8)     # Q30..Q40
    low = lowsel < lowq_per_65536
    qual[rows[low], lowpos[low]] = U64(33).astype(np.uint8) + lowq[low]                 # Q2..Q28
    out[:, 17 + 2 * L] = 10
    return out.reshape(-1)


# ---------------------------------------------------------------------------------------------------
# shapes 1-3 of K0 (csrc/synth_gen.h): Bar-seq, dual fixed, dual delimiters — the workloads of BASELINE configs 4 and 5
# ---------------------------------------------------------------------------------------------------
SY1_SUB, SY1_LACK, SY1_LEN = 3277, 5243, 6554
SY2_XMUT, SY2_YMUT, SY2_XRAND, SY2_MISPAIR = 6554, 9830, 11796, 6554
F_EXTRA, F_BG0 = 5, 39

BARSEQ_US, BARSEQ_DS = b"GTTCAGAGTTCT", b"CTGAATAGGCCA"
DUAL_DELIMS = (b"ACCGGT", b"TTGACA", b"GGATCC", b"CAATTG")          # U1, D1, U2, D2


def shape_spec(config: str):
    """generator spec of the non-guide workloads: '4' Bar-seq (config 4), '5a' dual fixed, '5b' dual delimiters"""
    base_ = dict(read_len=75, feat_len=20, cum_exact=0, cum_sub1=0, cum_sub2=0, cum_sub3=0, cum_n=0)
    if config == "4":
        return dict(base_, seed=4, shape=1, delims=(BARSEQ_US, BARSEQ_DS), lowq_per_65536=int(round(0.08 * 65536)))
    if config == "5a":
        return dict(base_, seed=5, shape=2, delims=(), lowq_per_65536=int(round(0.08 * 65536)))
    if config == "5b":
        return dict(base_, seed=6, shape=3, delims=DUAL_DELIMS, lowq_per_65536=int(round(0.08 * 65536)))
    raise ValueError(config)


def dual_keys(n_pairs: int, seed: int = 5, length: int = 20):
    """config-5 library: n_pairs 'X:Y' keys plus their single X and Y entries (names d/x/y%06d); returns
    (names, keys, xs, ys); the generator takes xs + ys as its guide table"""
    _, seqs = make_library(seed ^ 0xD0A1, 2 * n_pairs, length)
    xs, ys = seqs[:n_pairs], seqs[n_pairs:]
    names, keys = [], []
    for k in range(n_pairs):
        names += ["d%06d" % k, "x%06d" % k, "y%06d" % k]
        keys += [xs[k] + b":" + ys[k], xs[k], ys[k]]
    return names, keys, xs, ys


def shaped_reads(guides, first_read: int, n_reads: int, *, seed, read_len, feat_len, shape, delims=(), lowq_per_65536, **_):
    """reads [first_read, first_read + n_reads) of shapes 1-3 -> uint8[n_reads * (2L+18)]; numpy restatement of
    synth_record (csrc/synth_gen.h), bit-identical (tests/test_synth_shapes.py, tests/test_gpu_parity.py)"""
    L, F = int(read_len), int(feat_len)
    assert shape in (1, 2, 3) and 4 <= F <= 32 and F <= L <= 150
    G = np.frombuffer(b"".join(guides), dtype=np.uint8).reshape(len(guides), F)
    n_g = len(guides) if shape == 1 else len(guides) // 2
    code_of = np.zeros(256, dtype=np.int64)
    for c, ch in enumerate(b"ACGT"):
        code_of[ch] = c
    n = int(n_reads)
    idx = np.arange(first_read, first_read + n, dtype=np.uint64)
    b = base(seed, idx)
    r0, r1, r2 = rnd(b, F_MISC), rnd(b, F_GUIDE), rnd(b, F_EXTRA)
    cls = (r0 & U64(0xFFFF)).astype(np.int64)
    lowsel = ((r0 >> U64(16)) & U64(0xFFFF)).astype(np.int64)
    lowpos = ((((r0 >> U64(32)) & U64(0xFFFF)) * U64(L)) >> U64(16)).astype(np.int64)
    lowq = (U64(2) + ((((r0 >> U64(48)) & U64(0xFFFF)) * U64(27)) >> U64(16))).astype(np.uint8)
    rec = 2 * L + 18
    out = np.empty((n, rec), dtype=np.uint8)
    out[:, 0] = ord("@")
    out[:, 1] = ord("S")
    v = idx.copy()
    for d in range(11):
        out[:, 12 - d] = (v % U64(10)).astype(np.uint8) + ord("0")
        v //= U64(10)
    out[:, 13] = 10
    seq = out[:, 14:14 + L]
    for w in range((L + 31) // 32):
        bg = rnd(b, F_BG0 + w)
        for j in range(32):
            if w * 32 + j >= L:
                break
            seq[:, w * 32 + j] = ACGT[((bg >> U64(2 * j)) & U64(3)).astype(np.int64)]
    rows = np.arange(n)
    p = np.zeros(n, dtype=np.int64)
    NONE = np.full(n, -1, dtype=np.int64)

    def put(src, length, mpos, mdelta, active=None):
        """src: (n, >= length) per-read bytes or one bytes object; advances p by length for the active reads"""
        nonlocal p
        if isinstance(src, (bytes, bytearray)):
            src = np.broadcast_to(np.frombuffer(bytes(src), dtype=np.uint8), (n, len(src)))
        act = np.ones(n, dtype=bool) if active is None else active
        for k in range(length):
            m = act & (p + k < L)
            c = src[:, k].copy()
            mut = m & (mpos == k)
            c[mut] = ACGT[(code_of[c[mut]] + 1 + mdelta[mut] % 3) % 4]
            seq[rows[m], (p + k)[m]] = c[m]
        p = p + np.where(act, length, 0)

    if shape == 1:
        u32 = r1 & U64(0xFFFFFFFF)
        sq = (u32 * u32) >> U64(32)
        bi = ((sq * U64(n_g)) >> U64(32)).astype(np.int64)
        which = ((r2 >> U64(3)) & U64(1)).astype(np.int64)
        mdelta = ((r2 >> U64(16)) & U64(3)).astype(np.int64)
        sub, lack, odd = cls < SY1_SUB, (cls >= SY1_SUB) & (cls < SY1_LACK), (cls >= SY1_LACK) & (cls < SY1_LEN)
        us, ds = delims
        sel8 = ((r2 >> U64(8)) & U64(0xFF)).astype(np.int64)
        p = (r2 & U64(7)).astype(np.int64)
        skip = lack & (which == 0)
        put(us, len(us), np.where(sub & (which == 0), (sel8 * len(us)) >> 8, NONE), mdelta, ~skip)
        p = p + np.where(skip, len(us), 0)
        short = odd & (((r2 >> U64(4)) & U64(1)) == U64(1))
        longer = odd & ~short
        bc = G[bi]
        put(bc, F - 1, NONE, mdelta)
        put(bc[:, F - 1:], 1, NONE, mdelta, ~short)
        p = p + np.where(longer, 1, 0)
        skip = lack & (which == 1)
        put(ds, len(ds), np.where(sub & (which == 1), (sel8 * len(ds)) >> 8, NONE), mdelta, ~skip)
        p = p + np.where(skip, len(ds), 0)
    else:
        k = (((r1 & U64(0xFFFFFFFF)) * U64(n_g)) >> U64(32)).astype(np.int64)
        other = (((r2 >> U64(32)) * U64(n_g)) >> U64(32)).astype(np.int64)
        ky = np.where((r2 & U64(0xFFFF)).astype(np.int64) < SY2_MISPAIR, other, k)
        X, Y = G[k], G[n_g + ky]
        mpos = ((((r2 >> U64(16)) & U64(0xFF)) * U64(F)) >> U64(8)).astype(np.int64)
        mdelta = ((r2 >> U64(24)) & U64(3)).astype(np.int64)
        xmut, ymut, xrand = cls < SY2_XMUT, (cls >= SY2_XMUT) & (cls < SY2_YMUT), (cls >= SY2_YMUT) & (cls < SY2_XRAND)
        if shape == 2:
            put(X, F, np.where(xmut, mpos, NONE), mdelta, ~xrand)
            p = p + np.where(xrand, F, 0) + 10
            put(Y, F, np.where(ymut, mpos, NONE), mdelta)
        else:
            u1, d1, u2, d2 = delims
            p = ((r2 >> U64(26)) & U64(3)).astype(np.int64)
            put(u1, len(u1), NONE, mdelta)
            put(X, F, np.where(xmut, mpos, NONE), mdelta, ~xrand)
            p = p + np.where(xrand, F, 0)
            put(d1, len(d1), NONE, mdelta)
            p = p + (((r2 >> U64(28)) & U64(3)) % U64(3)).astype(np.int64)
            put(u2, len(u2), NONE, mdelta)
            put(Y, F, np.where(ymut, mpos, NONE), mdelta)
            put(d2, len(d2), NONE, mdelta)
    out[:, 14 + L] = 10
    out[:, 15 + L] = ord("+")
    out[:, 16 + L] = 10
    qual = out[:, 17 + L:17 + 2 * L]
    for w in range((L + 7) // 8):
        rq = rnd(b, F_QUAL0 + w)
        for j in range(8):
            kk = w * 8 + j
            if kk >= L:
                break
            byte = (rq >> U64(8 * j)) & U64(0xFF)
            qual[:, kk] = (U64(63) + ((byte * U64(11)) >> U64(8))).astype(np.uint8)
    low = lowsel < lowq_per_65536
    qual[rows[low], lowpos[low]] = U64(33).astype(np.uint8) + lowq[low]
    out[:, 17 + 2 * L] = 10
    return out.reshape(-1)


# ---------------------------------------------------------------------------------------------------
# small scalar RNG for the irregular workloads (bar-seq / dual feature / fuzz); stream = (seed, counter)
# ---------------------------------------------------------------------------------------------------
class SM64:
    def __init__(self, seed: int):
        self.s = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.c = 0

    def u64(self) -> int:
        self.c += 1
        z = (self.s + self.c * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.u64() % n if n > 0 else 0

    def chance(self, p: float) -> bool:
        return (self.u64() >> 11) * (1.0 / (1 << 53)) < p

    def choice(self, seq):
        return seq[self.below(len(seq))]

    def dna(self, n: int, alphabet=b"ACGT") -> bytes:
        return bytes(alphabet[self.below(len(alphabet))] for _ in range(n))


def mutate(rng: SM64, s: bytes, k: int, alphabet=b"ACGT") -> bytes:
    """k substitutions at distinct positions"""
    s = bytearray(s)
    pos = list(range(len(s)))
    for _ in range(min(k, len(s))):
        p = pos.pop(rng.below(len(pos)))
        opts = [c for c in alphabet if c != s[p]]
        s[p] = rng.choice(opts)
    return bytes(s)


def qual_line(rng: SM64, n: int, p_low: float = 0.08) -> bytes:
    q = bytearray(63 + rng.below(11) for _ in range(n))
    if n and rng.chance(p_low):
        q[rng.below(n)] = 33 + 2 + rng.below(27)
    return bytes(q)


def barseq_reads(seed: int, n_reads: int, *, read_len=75, us=b"GTTCAGAGTTCT", ds=b"CTGAATAGGCCA", n_barcodes=2000):
    """config-4 shaped Bar-seq reads: stagger + US + barcode(20, sometimes 19/21) + DS + pad"""
    rng = SM64(seed * 1000003 + 4)
    pool = [rng.dna(20) for _ in range(n_barcodes)]
    out = []
    for i in range(n_reads):
        # heavy-tailed abundance: square of a uniform picks low indices more often
        u = rng.below(1 << 20) / float(1 << 20)
        bc = pool[int(u * u * n_barcodes)]
        t = rng.below(100)
        if t < 2:
            bc = bc[:19] if rng.chance(0.5) else bc + rng.dna(1)
        u_, d_ = us, ds
        t = rng.below(100)
        if t < 5:
            if rng.chance(0.5):
                u_ = mutate(rng, us, 1)
            else:
                d_ = mutate(rng, ds, 1)
        elif t < 8:
            if rng.chance(0.5):
                u_ = rng.dna(len(us))
            else:
                d_ = rng.dna(len(ds))
        s = rng.dna(rng.below(8)) + u_ + bc + d_
        s = (s + rng.dna(read_len))[:read_len]
        out.append(b"@B%010d\n" % i + s + b"\n+\n" + qual_line(rng, read_len) + b"\n")
    return b"".join(out)


def dual_library(seed: int, n_pairs: int):
    """config-5 library: n_pairs 'X:Y' dual keys plus their single X and Y entries"""
    rng = SM64(seed * 7919 + 5)
    xs = [rng.dna(20) for _ in range(n_pairs)]
    ys = [rng.dna(20) for _ in range(n_pairs)]
    keys, names, seen = [], [], set()
    for k, (x, y) in enumerate(zip(xs, ys)):
        for nm, s in (("d%05d" % k, x + b":" + y), ("x%05d" % k, x), ("y%05d" % k, y)):
            if s not in seen:
                seen.add(s)
                keys.append(s)
                names.append(nm)
    return names, keys, xs, ys


def dual_reads(seed: int, n_reads: int, xs, ys, *, read_len=75, mode="fixed", u1=b"ACCGGT", d1=b"TTGACA", u2=b"GGATCC",
               d2=b"CAATTG"):
    """config-5 shaped reads.  fixed: X at 0, Y at 30.  delim: u1 X d1 ... u2 Y d2"""
    rng = SM64(seed * 104729 + (5 if mode == "fixed" else 6))
    out = []
    for i in range(n_reads):
        k = rng.below(len(xs))
        x, y = xs[k], ys[k if rng.below(100) < 90 else rng.below(len(ys))]
        t = rng.below(100)
        if t < 10:
            x = mutate(rng, x, 1)
        elif t < 15:
            y = mutate(rng, y, 1)
        elif t < 18:
            x = rng.dna(20)
        if mode == "fixed":
            s = x + rng.dna(10) + y
        else:
            s = rng.dna(rng.below(4)) + u1 + x + d1 + rng.dna(rng.below(3)) + u2 + y + d2
        s = (s + rng.dna(read_len))[:read_len]
        out.append(b"@D%010d\n" % i + s + b"\n+\n" + qual_line(rng, read_len, 0.15) + b"\n")
    return b"".join(out)
