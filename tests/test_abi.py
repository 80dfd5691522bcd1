"""
The drop-in boundary without a GPU: libf2q.so loads, exports exactly the entry points include/f2q.h declares, the ctypes
binding covers all of them, and the library refuses to work without a device instead of falling back to the CPU.
"""
import ctypes as C
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f2q = importlib.import_module("2fast2q_b200")
lib = f2q._lib


def declared():
    text = open(os.path.join(ROOT, "include", "f2q.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int|void|uint64_t)\s+\*?\s*(f2q_[a-z0-9_]+)\s*\(", text, flags=re.M)
    assert len(names) >= 25 and len(set(names)) == len(names)
    return names


def test_library_exports_every_declared_symbol():
    names = declared()
    L = C.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/f2q.h but not exported by libf2q.so"


def test_binding_covers_the_header_exactly():
    assert set(lib.SYMBOLS) == set(declared())


def test_exports_nothing_else_with_the_prefix():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("f2q_")}
    assert exported == set(declared())


def test_no_cpu_fallback_without_a_device():
    L = lib.load()
    assert L.f2q_abi_version() == lib.ABI_VERSION
    if L.f2q_device_count() > 0:
        pytest.skip("a B200 is visible: the refusal path needs a machine without one")
    cfg = lib.make_config()
    with pytest.raises(lib.F2QError) as e:
        lib.Engine(cfg, 0)
    assert e.value.code == -7 and "no CPU path" in str(e.value)          # F2Q_ENODEVICE
    with pytest.raises(lib.F2QError):
        lib.border_finder_device(b"ACGT", b"TTACGTTT", 0)
