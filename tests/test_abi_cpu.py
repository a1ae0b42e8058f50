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


def test_widening_entry_points_validate_arguments(lib):
    """rxb_jpeg_decode_gray / rxb_load_norm_affine: sizes and pointers are checked before any device work."""
    assert lib.rxb_jpeg_decode_workspace_bytes(768, 512, 512) == 768 * 64 * 64 * 128     # 128 B per 8x8 block
    assert lib.rxb_jpeg_decode_workspace_bytes(1, 37, 53) == 5 * 7 * 128                 # partial blocks count
    assert lib.rxb_jpeg_decode_workspace_bytes(0, 512, 512) == 0
    assert lib.rxb_jpeg_decode_gray(None, None, None, 0, 512, 512, None, None, None, 0, None) == 0   # empty batch
    assert lib.rxb_jpeg_decode_gray(None, None, None, 3, 512, 512, None, None, None, 0, None) == -1
    assert b"null" in lib.rxb_last_error()
    assert lib.rxb_jpeg_decode_gray(None, None, None, -1, 512, 512, None, None, None, 0, None) == -1
    buf = (ctypes.c_uint8 * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.rxb_jpeg_decode_gray(p, p, p, 1, 0, 512, p, p, None, 0, None) == -1
    assert b"size" in lib.rxb_last_error()
    assert lib.rxb_jpeg_decode_gray(p, p, p, 1, 512, 512, p, p, p, 16, None) == -1       # workspace too small
    assert b"workspace" in lib.rxb_last_error()
    assert lib.rxb_load_norm_affine(None, 1, 512, 512, None, None, None, None, None, None, None, 1, None, 0, 512, 512, 0,
                                    None) == 0                                           # empty batch
    assert lib.rxb_load_norm_affine(None, 1, 512, 512, None, None, None, None, None, None, None, 1, None, 2, 512, 512, 0,
                                    None) == -1
    assert lib.rxb_load_norm_affine(p, 1, 32, 32, p, p, p, p, p, p, p, 1, p, 1, 48, 48, 0, None) == -1   # crop > image
    assert b"crop" in lib.rxb_last_error()
    assert lib.rxb_load_norm_affine(p, 1, 32, 32, p, p, p, p, p, p, p, 1, p, 1, 31, 31, 2, None) == -1   # S2D32 odd
    assert lib.rxb_load_norm_affine(p, 1, 32, 32, p, p, p, p, p, p, p, 1, p, 1, 32, 32, 7, None) == -1   # bad format


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_widening_ops_fail_loudly_without_gpu():
    from recursion_cellular_image_classification_b200 import ops
    blob, offsets = ops.pack_jpeg_buffers([b"\xff\xd8\xff\xd9"])
    with pytest.raises(_lib.RxbError):
        ops.jpeg_decode_gray(blob, offsets, (8, 8))
    z = torch.zeros(1, dtype=torch.int32)
    with pytest.raises(_lib.RxbError):
        ops.load_norm_affine(torch.zeros(1, 6, 8, 8, dtype=torch.uint8), z, z, torch.zeros(1, dtype=torch.uint8),
                             torch.zeros(1, 2, 3, dtype=torch.float64), torch.zeros(1, 2, dtype=torch.int32),
                             torch.zeros(1, 6), torch.ones(1, 6), (8, 8), ops.OUT_F32_NCHW)


def test_header_is_plain_c_and_the_library_links_from_c(lib, tmp_path):
    """include/rxb.h compiles as C99 with -pedantic and a C program linked against librxb.so gets answers from the
    host-side entry points (version, planning queries, argument validation)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cabi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "cabi_smoke.c"), "-o", exe, "-L", libdir, "-lrxb",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "version 100" in out and "params 8098964" in out and "jpeg_ws 402653184" in out
    assert "rc -1: rxb_stats_accumulate: null pointer" in out
