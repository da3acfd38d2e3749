import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def _make(directory, target):
    path = os.path.join(ROOT, directory, target)
    if not os.path.exists(path) and os.path.exists("/usr/bin/make"):
        subprocess.run(["make", "-C", os.path.join(ROOT, directory)], check=True, capture_output=True)
    return path


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (checker).  Only tests may load it."""
    from katana_jl_b200.binding import KtnLibrary
    return KtnLibrary(_make("oracle", "libktn_oracle.so"))


@pytest.fixture(scope="session")
def emu_lib():
    """Host emulator of the CUDA kernels over the real compiler output (tests/emu)."""
    from katana_jl_b200.binding import KtnLibrary
    return KtnLibrary(_make("tests/emu", "libktn_emu.so"))


@pytest.fixture(scope="session")
def cuda_lib():
    from katana_jl_b200.binding import load_cuda_library
    return load_cuda_library()


@pytest.fixture(scope="session")
def synth_lib():
    """libktn_synth.so: the synthetic generators alone (no device call, not the product library)."""
    from katana_jl_b200.binding import SynthLibrary
    _make("katana.jl_b200/csrc", "../libktn_synth.so")
    return SynthLibrary()
