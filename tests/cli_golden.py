"""loader of tests/golden/cli_cases.json.gz (made by tests/golden/make_cli_golden.py from the unmodified reference)"""
import base64
import gzip
import json
import os
import re

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cli_cases.json.gz")
_cache = None


def load():
    global _cache
    if _cache is None:
        with gzip.open(_PATH, "rb") as f:
            _cache = json.loads(f.read().decode())
    return _cache


def case(name):
    return [c for c in load()["cli"] if c["name"] == name][0]


def write_inputs(c, root):
    """creates <root>/in/* (+ <root>/library.csv); returns the reference-style argv"""
    src, out = os.path.join(root, "in"), os.path.join(root, "out")
    os.makedirs(src, exist_ok=True)
    os.makedirs(out, exist_ok=True)
    for fn, b64 in c["files"].items():
        with open(os.path.join(src, fn), "wb") as f:
            f.write(base64.b64decode(b64))
    argv = ["-c", "--s", src, "--o", out] + list(c["args"])
    if c["library"] is not None:
        with open(os.path.join(root, "library.csv"), "w", newline="") as f:
            f.write(c["library"])
        argv += ["--g", os.path.join(root, "library.csv")]
    return argv


_TIME = re.compile(r"ran in \S+ \S+ for file")


def mask_reads_csv(text):
    """the statistics sentence without its wall-clock part"""
    return _TIME.sub("ran in <T> for file", text)


def mask_stats_csv(text, tmp=None):
    """compiled_stats.csv without run times and without the '#cmd used' line (paths / added flags differ)"""
    out = []
    for line in text.split("\r\n"):
        if line.startswith("#cmd used") or line.startswith('"#cmd used'):
            continue
        cols = line.split(",")
        if not line.startswith("#") and not line.startswith('"#') and len(cols) == 9:
            cols[1], cols[2] = "<T>", "<U>"
            line = ",".join(cols)
        out.append(line)
    return "\r\n".join(out)
