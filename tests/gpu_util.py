"""shared helpers of the -m gpu tests: run a golden case through libf2q (C-ABI via ctypes)"""
import importlib

f2q = importlib.import_module("2fast2q_b200")
lib = f2q._lib


def run_case(params, library, fastq, chunk=None, **options):
    """returns (counts list | ec dict, stats dict) computed on the GPU"""
    cfg = lib.make_config(**params)
    with lib.Engine(cfg, 0, None, **options) as e:
        if library is not None:
            e.set_library([s for _, s in library])
        counts, stats = e.run(fastq, chunk)
        if library is not None:
            return [int(x) for x in counts], stats
        return e.ec_items(), stats


def check_case(c, fastq, chunk=None, **options):
    got, stats = run_case(c["params"], c.get("library"), fastq, chunk, **options)
    tag = (c["name"], chunk, options)
    assert stats == c["stats"], tag
    if "counts" in c:
        assert got == c["counts"], tag
    else:
        assert got == {k.encode("latin-1"): v for k, v in c["ec"]}, tag
