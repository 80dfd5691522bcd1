"""
Synthetic FASTQ generators for the tests: re-exported from 2fast2q_b200/synth.py (the numpy restatement of the CUDA
generator K0 lives beside the library that holds K0, so that bench.py's measured legs import nothing from oracle/).
"""
import importlib as _importlib

_m = _importlib.import_module("2fast2q_b200.synth")
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
