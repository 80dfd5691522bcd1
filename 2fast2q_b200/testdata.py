"""
Test-mode data (`2fast2q -c -t`, fast2q.py:1237-1240).

The reference bundles fast2q/data/example.fastq.gz + D39V_guides.csv; the FASTQ is not part of the reference checkout
(`.MISSING_LARGE_BLOBS`), so this package ships a SURROGATE: a deterministic FASTQ constructed so that the reference's
answer on it is exactly the reference's own tests/compiled.csv (kept here as data/expected_compiled.csv).  The surrogate
was validated against the unmodified reference when the golden fixtures were made (tests/golden/make_golden.py); its
sha256 is pinned in tests/golden/config1_surrogate.json.  data/example.fastq.gz is written on first use (git-ignored).
"""
from __future__ import annotations

import gzip
import os

from . import synth

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
GUIDES = os.path.join(DATA, "D39V_guides.csv")
EXPECTED = os.path.join(DATA, "expected_compiled.csv")
EXAMPLE = os.path.join(DATA, "example.fastq.gz")


def load_guides_csv(path):
    """features_loader semantics for a clean comma file (fast2q.py:148-166): first sequence wins"""
    lib, seen = [], set()
    with open(path) as f:
        for line in f:
            parts = line.rstrip().split(",")
            seq = parts[1].upper().replace(" ", "")
            if seq not in seen:
                seen.add(seq)
                lib.append((parts[0], seq))
    return lib


def config1_surrogate(guides_csv, compiled_csv):
    import numpy as np
    lib = load_guides_csv(guides_csv)
    want = {}
    with open(compiled_csv) as f:
        for line in f:
            if line.startswith("#"):
                continue
            n, c = line.rstrip().split(",")
            want[n] = int(c)
    seqs = [s.encode() for _, s in lib]
    arr = np.frombuffer(b"".join(seqs), dtype=np.uint8).reshape(len(seqs), 20)
    r = synth.SM64(0xC0FF1)
    recs = []
    for gi, (name, seq) in enumerate(lib):
        s = seq.encode()
        for _ in range(want.get(name, 0)):
            v = s
            if r.below(100) < 15:
                for _try in range(20):
                    cand = synth.mutate(r, s, 1)
                    d = (arr != np.frombuffer(cand, dtype=np.uint8)).sum(axis=1)
                    if (d <= 1).sum() == 1:        # unique within distance 1 -> the reference assigns it to gi
                        v = cand
                        break
            recs.append((v + r.dna(30), synth.qual_line(r, 50, 0.0)))
    n_al = len(recs)
    for _ in range(n_al // 12):                     # low quality inside the window -> quality_failed
        g = r.choice(seqs)
        q = bytearray(synth.qual_line(r, 50, 0.0))
        q[r.below(20)] = 33 + r.below(29)
        recs.append((g + r.dna(30), bytes(q)))
    for _ in range(n_al // 20):                     # unalignable
        for _try in range(50):
            cand = r.dna(20)
            d = (arr != np.frombuffer(cand, dtype=np.uint8)).sum(axis=1)
            if d.min() > 1:
                break
        recs.append((cand + r.dna(30), synth.qual_line(r, 50, 0.0)))
    # deterministic shuffle
    order = list(range(len(recs)))
    for i in range(len(order) - 1, 0, -1):
        j = r.below(i + 1)
        order[i], order[j] = order[j], order[i]
    data = b"".join(b"@E%07d\n" % k + recs[o][0] + b"\n+\n" + recs[o][1] + b"\n" for k, o in enumerate(order))
    return lib, want, data


def ensure_example() -> str:
    """path of the surrogate example.fastq.gz, generated on first use"""
    if not os.path.exists(EXAMPLE):
        _, _, data = config1_surrogate(GUIDES, EXPECTED)
        tmp = EXAMPLE + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as raw, gzip.GzipFile(filename="", mode="wb", fileobj=raw, compresslevel=6, mtime=0) as f:
            f.write(data)
        os.replace(tmp, EXAMPLE)
    return EXAMPLE
