"""CPU-side ABI conformance: the C library loads, exports every symbol the header declares and
behaves sanely without a GPU (no compute calls)."""
import ctypes
import json
import os
import re
import subprocess

import pytest

import helpers as H  # noqa: F401  (sets sys.path)
from beast_mcmc_b200 import beagle, build

ROOT = H.ROOT


@pytest.fixture(scope="module")
def lib():
    build.build_engine()
    return beagle.load_library()


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "libhmsbeagle_b200.h")).read()
    return re.findall(r"BEAGLE_DLLEXPORT\s+[\w\s\*]+?\b(\w+)\s*\(", txt)


def test_header_declares_what_python_binds():
    assert sorted(set(_header_functions())) == sorted(beagle.exported_symbols())


def test_every_declared_symbol_is_exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", build.lib_path()], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [f for f in _header_functions() if f not in exported]
    assert not missing, missing


def test_version_gate(lib):
    v = lib.beagleGetVersion().decode()
    m = re.match(r"(\d+)\.(\d+)\.(\d+).*", v)       # BeagleInfo regex; BeastMain.java:945-950 needs major >= 4
    assert m and int(m.group(1)) >= 4


def test_resource_zero_is_cpu(lib):
    res = beagle.BeagleFactory.getResourceDetails()
    assert len(res) >= 1 and "CPU" in res[0].name


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(beagle.BeagleException) as e:
        beagle.BeagleFactory.loadBeagleInstance(3, 5, 3, 4, 10, 2, 10, 1, 6, [1, 0], 0, 0)
    assert e.value.errCode == beagle.BeagleErrorCode.NO_RESOURCE_ERROR


def test_uninitialized_instance_error(lib):
    assert lib.beagleFinalizeInstance(12345) == beagle.BeagleErrorCode.UNINITIALIZED_INSTANCE_ERROR
    assert lib.beagleResetScaleFactors(777, 0) == beagle.BeagleErrorCode.UNINITIALIZED_INSTANCE_ERROR


def test_jar_interface_is_covered():
    """Every method of the jar's beagle.Beagle interface exists on the Python mirror."""
    abi = json.load(open(os.path.join(ROOT, "tests", "golden", "beagle_jar_abi.json")))
    later = {"setRootPrePartials", "calculateCrossProductDifferentials", "calculateEdgeDerivative",
             "getSiteDerivatives", "calculateEdgeLogLikelihoods"}     # SURVEY.md 8f "next" rows
    names = {m["name"] for m in abi["beagle_interface"]}
    missing = [n for n in sorted(names - later) if not hasattr(beagle.BeagleJNIImpl, n)]
    assert not missing, missing
