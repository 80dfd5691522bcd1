"""loaders for the committed golden fixtures (tests/golden/*.json[.gz])"""
import gzip
import json
import os

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fix(c):
    if "fastq" in c:
        c["fastq"] = c["fastq"].encode("latin-1")
    return c


def kat():
    return [_fix(c) for c in json.load(open(os.path.join(HERE, "kat_cases.json")))]


def fuzz():
    with gzip.open(os.path.join(HERE, "fuzz_cases.json.gz"), "rb") as f:
        return [_fix(c) for c in json.loads(f.read().decode())]


def shaped():
    return json.load(open(os.path.join(HERE, "shaped_cases.json")))


def primitives():
    return json.load(open(os.path.join(HERE, "primitive_cases.json")))


def config1():
    return json.load(open(os.path.join(HERE, "config1_surrogate.json")))
