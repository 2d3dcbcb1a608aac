"""CPU tests of the boundary: the C-ABI library loads and exports every symbol the header
declares; host-side mirrors behave like the reference interface.  No compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "cloudbrush_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from cloudbrush_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/cloudbrush_gpu.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms  # the ctypes binding knows them all


def test_default_config_matches_brushconfig_defaults():
    from cloudbrush_b200 import _lib
    lib = _lib.load()
    c = _lib.cb_config()
    lib.cb_default_config(C.byref(c))
    assert (c.abi_version, c.up_kmer, c.low_kmer) == (1, 2000, 1)        # BrushConfig.java:83-84
    assert abs(c.majority - 0.6) < 1e-7 and abs(c.pwm_n - 0.1) < 1e-7   # BrushConfig.java:62-63
    assert lib.cb_version().startswith(b"cloudbrush-b200")
    assert C.sizeof(_lib.cb_edge) == 12 and C.sizeof(_lib.cb_node) == 8


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cloudbrush_b200 import BrushConfig, BrushGpu, BrushGpuError
    with pytest.raises(BrushGpuError) as ei:
        BrushGpu(BrushConfig())
    assert ei.value.code == -2  # CB_E_NODEVICE


def test_null_and_bad_arguments_return_codes():
    from cloudbrush_b200 import _lib
    lib = _lib.load()
    assert lib.cb_create(None, None) == -1
    c = _lib.cb_config()
    lib.cb_default_config(C.byref(c))
    c.abi_version = 99
    h = C.c_void_p()
    assert lib.cb_create(C.byref(c), C.byref(h)) == -1
    assert lib.cb_preprocess(None) == -1
    v = C.c_int64()
    assert lib.cb_get_counter(None, b"nodes", C.byref(v)) == -1


def test_native_nccl_entry_points_reject_bad_arguments():
    """cb_nccl_get_unique_id / cb_init_nccl validate before touching NCCL (no GPU needed)."""
    from cloudbrush_b200 import _lib
    lib = _lib.load()
    assert lib.cb_nccl_get_unique_id(None) == -1
    assert lib.cb_init_nccl(None, None) == -1
    txt = open(os.path.join(ROOT, "include", "cloudbrush_gpu.h")).read()
    assert "CB_NCCL_ID_BYTES 128" in txt and "allreduce_max_u8_dev" in txt


def test_running_job_lookalike():
    from cloudbrush_b200.stages import RunningJob
    j = RunningJob("preprocess", {"nodes": 5, "reads_good": 7}, 0.1)
    assert j.isSuccessful() and j.getJobID().startswith("job_gpu_")
    assert j.getCounters().findCounter("Brush", "nodes").getValue() == 5
    assert j.getCounters().findCounter("Brush", "never_incremented").getValue() == 0  # Hadoop semantics
    assert j.getCounters().findCounter("Other", "nodes").getValue() == 0


def test_synthetic_generator_is_deterministic_and_well_formed():
    from cloudbrush_b200 import synth
    a, ma = synth.make_config("cfg3_50k", scale_genome=5000)
    b, mb = synth.make_config("cfg3_50k", scale_genome=5000)
    assert np.array_equal(a, b) and ma == mb
    lines = bytes(a).split(b"\n")
    assert lines[-1] == b"" and len(lines) - 1 == ma["n_reads"] == round(5000 * 50 / 100)
    ids = [l.split(b"\t")[0] for l in lines[:-1]]
    assert ids[0] == b"1" and ids[9] == b"A" and ids[15] == b"10" and ids[-1] == format(ma["n_reads"], "X").encode()
    assert all(len(l.split(b"\t")[1]) == 100 and set(l.split(b"\t")[1]) <= set(b"ACGT") for l in lines[:-1])
    # forward-only config: every read is a substring of one genome
    c, mc = synth.make_config("cfg2_50k", scale_genome=3000)
    assert mc["both_strands"] is False


def test_bucket_plan_sizes_buckets_for_the_bucket_kernel():
    """make_bucket_plan (host logic, no device): digit widths from {3, 6, 8, 9, 10}, enough buckets that the
    average stays at or below 640 tuples, two passes refused beyond 2^20 buckets of 640."""
    from cloudbrush_b200 import _lib
    lib = _lib.load()
    b1, b2, nb, fits = C.c_int32(), C.c_int32(), C.c_uint32(), C.c_int32()
    for n in (0, 1, 1000, 264_112, 69_014_800, 126_674_380, 344_000_000, 671_088_640, 671_088_641, 2_754_195_570):
        assert lib.cb_debug_bucket_plan(n, C.byref(b1), C.byref(b2), C.byref(nb), C.byref(fits)) == 0
        assert b1.value in (3, 6, 8, 9, 10) and b2.value in (3, 6, 8, 9, 10) and b1.value >= b2.value
        assert nb.value == 1 << (b1.value + b2.value)
        if fits.value:
            assert n <= 640 * nb.value
        assert bool(fits.value) == (n <= 640 << 20)
    assert lib.cb_debug_bucket_plan(69_014_800, C.byref(b1), C.byref(b2), C.byref(nb), C.byref(fits)) == 0
    assert (b1.value, b2.value) == (9, 8)   # config 2: 131 072 buckets of 526 tuples
