"""Link hygiene of the native module, checked WITHOUT a GPU.

Round 1's GPU suite died with SIGSEGV because the extension carried a static libstdc++ next to torch's dynamic one
(any ``std::ostringstream << int`` crashed).  These tests fail on the CPU box if that ever comes back."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _so():
    sos = glob.glob(os.path.join(ROOT, "shallowspeed_b200", "_C*.so"))
    if not sos:
        pytest.skip("native module not built")
    return sos[0]


def test_dt_needed_lists_dynamic_libstdcxx_and_nccl():
    dyn = subprocess.run(["readelf", "-d", _so()], capture_output=True, text=True, check=True).stdout
    assert "libstdc++.so.6" in dyn, "libstdc++ linked statically"
    assert "libnccl.so.2" in dyn, "NCCL not linked"


def test_no_iostream_definitions_exported():
    syms = subprocess.run(["nm", "-D", "--defined-only", _so()], capture_output=True, text=True, check=True).stdout
    bad = [l for l in syms.splitlines() if "_ZNSo3putEc" in l or "_ZNSolsEi" in l or "_ZNSt8ios_base4InitC1Ev" in l]
    assert not bad, bad


def test_stream_formatting_and_exceptions_work_next_to_torch():
    import torch  # noqa: F401  (loads the dynamic libstdc++ first, as every real process does)

    _so()
    from shallowspeed_b200 import _C

    assert _C.selftest_format(42) == "v=42 f=1.5 h=ff"
    assert _C.selftest_format(-7).startswith("v=-7 ")
