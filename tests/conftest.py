"""pytest configuration: the `gpu` marker, import paths, shared fixtures.

`-m "not gpu"` runs here (CPU only): oracle vs golden vectors, host logic, C-ABI symbol checks.
`-m gpu` runs on a B200: parity of the CUDA path (through the C ABI) against the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "fp-mash_b200", "py"), os.path.join(ROOT, "oracle"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle_py import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref/libmashref.so not built (reference tree absent)")
    return RefLib()


@pytest.fixture(scope="session")
def fpm():
    import fpmash_b200
    return fpmash_b200


@pytest.fixture(scope="session")
def ctx(fpm):
    # No skip-on-missing-GPU here: a gpu-marked test on a box without a device must fail
    # loudly (the product has no CPU fallback).
    c = fpm.Context(0)
    yield c
    c.close()
