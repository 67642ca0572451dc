"""CPU: the C-ABI library loads and exports every symbol include/shortseq_b200.h declares.
No compute calls are made (there is no GPU here); the product must fail loudly instead."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "shortseq_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssq_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    names = declared_symbols()
    for must in ("ssq_pack64", "ssq_pack192", "ssq_packvar", "ssq_decode64", "ssq_decode192", "ssq_decodevar",
                 "ssq_hamming_pairs64", "ssq_hamming_pairs192", "ssq_hamming_pairsvar", "ssq_hamming_refset",
                 "ssq_counter_create", "ssq_counter_insert", "ssq_counter_merge", "ssq_counter_pack_count",
                 "ssq_counter_export", "ssq_host_pack_count", "ssq_ctx_create", "ssq_ctx_sync", "ssq_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from shortseq_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert set(_lib.PROTOTYPES) == set(declared_symbols()), "ctypes prototypes and header disagree"
    assert _lib.lib().ssq_abi_version() == 1


def test_header_is_plain_c():
    """The boundary is extern "C" with pointers and sizes only: it must compile as C."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "shortseq_b200.h"\nint main(void){ssq_report r; r.code = SSQ_OK; return r.code;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", src, "-o",
                               os.path.join(d, "t.o")])


def test_no_cpu_fallback():
    """Without a CUDA device every product path raises; nothing routes through the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import shortseq_b200 as sq
    for call in (lambda: sq.pack("ACGT"), lambda: sq.pack_batch([b"ACGT"]), lambda: sq.ShortSeqCounter([b"ACGT"]),
                 lambda: sq.DeviceCounter(0)):
        with pytest.raises(sq.LibraryError):
            call()
    h = ctypes.c_void_p()
    from shortseq_b200 import _lib
    assert _lib.lib().ssq_ctx_create(0, ctypes.byref(h)) == _lib.ERR_CUDA
    assert b"CUDA" in _lib.lib().ssq_last_error() or b"device" in _lib.lib().ssq_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "shortseq_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "ssq_oracle" not in text.replace(
                    "oracle/ssq_oracle.c", ""), f
