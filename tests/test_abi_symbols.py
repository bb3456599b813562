"""CPU: the C-ABI library is built, loads, and exports every symbol include/commet_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

import commet_b200
from commet_b200 import api, build

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    build.build_lib()
    return ctypes.CDLL(str(api.lib_path()))


def declared_symbols():
    text = (ROOT / "include" / "commet_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(commet_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds():
    assert declared_symbols() == sorted(api.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_abi_version_and_reference_constants(lib):
    api.load_library()
    assert lib.commet_abi_version() == 1
    # include/bloom_filter.h:73-76 and src/index_and_search.cpp:73
    assert commet_b200.filter_bytes(33) == 1 << 32 and commet_b200.filter_bytes(27) == 1 << 26
    assert commet_b200.max_kmer(33) == 1_000_000_000
    assert commet_b200.max_kmer(32) == 500_000_000
    assert commet_b200.max_kmer(20) == 122070
    assert commet_b200.max_kmer(27) == 15_625_000


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(commet_b200.CommetError, match="no CPU fallback"):
        commet_b200.Context(0)
