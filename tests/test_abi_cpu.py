"""CPU tests of the C-ABI boundary: the library builds, loads, and exports every declared symbol;
the product path refuses to run without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from recursion_cellular_image_classification_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_are_all_exported_and_bound(lib):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "rxb.h")).read()
    declared = set(re.findall(r"\b(rxb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rxb_stream_t"}
    assert declared, "no declarations parsed"
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, "declared in include/rxb.h but not exported: %s" % missing
    unbound = sorted(declared - set(_lib.SIGNATURES))
    assert not unbound, "declared but not bound in _lib.SIGNATURES: %s" % unbound


def test_version_and_error_string(lib):
    assert lib.rxb_version() >= 100
    assert isinstance(lib.rxb_last_error(), bytes)


def test_argument_validation_happens_before_any_device_work(lib):
    rc = lib.rxb_stats_accumulate(None, None, 1, 512, 512, 6, 0, 1, None, None, None, None)
    assert rc == -1
    assert b"null" in lib.rxb_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    from recursion_cellular_image_classification_b200 import ops
    with pytest.raises(_lib.RxbError):
        ops.stats_accumulate(torch.zeros(1, 6, 16, 16, dtype=torch.uint8), torch.zeros(1, dtype=torch.int32), 1)
