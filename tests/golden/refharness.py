"""
Runs the UNMODIFIED reference (/root/reference/fast2q/fast2q.py) in-process.  Only usable in the build
container (the GPU box has no /root/reference); used by make_golden.py to produce the committed fixtures
and by tests/test_reference_live.py (skipped when the reference is absent).
"""
import os
import sys
import tempfile

REF_ROOT = os.environ.get("F2Q_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "fast2q", "fast2q.py"))


def load():
    """import fast2q.fast2q from the reference tree with the colorama/matplotlib stubs on the path"""
    for p in (REF_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("PYTHONWARNINGS", "ignore")
    import warnings
    warnings.filterwarnings("ignore")
    import fast2q.fast2q as ref
    return ref


def make_param(mode="C", miss=1, phred=30, length=20, start="0", upstream=None, downstream=None,
               miss_search_up=0, miss_search_down=0, qual_up=30, qual_down=30):
    """`param` as input_parser + initializer build it (fast2q.py:1246-1309, 1112-1129)."""
    p = {
        "miss": int(miss), "phred": int(phred), "length": int(length), "start": str(start),
        "upstream": upstream, "downstream": downstream,
        "miss_search_up": int(miss_search_up), "miss_search_down": int(miss_search_down),
        "qual_up": int(qual_up), "qual_down": int(qual_down),
        "big_file_split": False, "cpu": 1, "Progress bar": False,
        "Running Mode": "EC" if "EC" in str(mode).upper() else "C",
    }
    quality_list = "".join(chr(q + 33) for q in range(94))
    for k in ("phred", "qual_up", "qual_down"):
        if int(p[k]) <= 0:
            p[k] = 1
    p["quality_set"] = set(quality_list[:int(p["phred"]) - 1])
    p["quality_set_up"] = set(quality_list[:int(p["qual_up"]) - 1])
    p["quality_set_down"] = set(quality_list[:int(p["qual_down"]) - 1])
    return p


def run_reads_counter(fastq_bytes, library, **params):
    """library: list of (name, seq) in file order (Counter) or None (EC).
    returns (counts: dict seq->count in dict order, stats dict)"""
    ref = load()
    param = make_param(**params)
    feats = {}
    if library is not None:
        for name, seq in library:
            if seq not in feats:
                feats[seq] = ref.Features(name, 0)
    with tempfile.NamedTemporaryFile(suffix=".fastq", delete=False) as f:
        f.write(fastq_bytes)
        path = f.name
    try:
        out = ref.reads_counter(0, path, feats, param, {"failed_reads": set(), "passed_reads": {}})
    finally:
        os.unlink(path)
    features, _, stats = out
    return {k: v.counts for k, v in features.items()}, {k: int(v) for k, v in stats.items()}
