"""Access to the UNMODIFIED reference built into oracle/_ref by build_ref.py.

TEST INFRASTRUCTURE ONLY.  `load()` returns the reference's `shortseq` module
(or None if oracle/_ref has not been built); `raw_words(obj)` reads the packed
words straight out of the reference object's memory (layouts:
short_seq_64.pxd:11-14, short_seq_192.pxd:11-14, short_seq_var.pxd:14-17).
"""
import ctypes
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_MOD = None


def available():
    return os.path.exists(os.path.join(REF_DIR, "shortseq", "__init__.py"))


def load():
    global _MOD
    if _MOD is None:
        if not available():
            return None
        if "shortseq" in sys.modules:
            _MOD = sys.modules["shortseq"]
        else:
            sys.path.insert(0, REF_DIR)
            try:
                _MOD = importlib.import_module("shortseq")
            finally:
                sys.path.remove(REF_DIR)
    return _MOD


def raw_words(obj):
    """Packed u64 words of a reference ShortSeq64/192/Var object, as Python ints."""
    sq = load()
    n = len(obj)
    if type(obj) is sq.ShortSeq64:
        return [ctypes.c_uint64.from_address(id(obj) + 16).value]
    if type(obj) is sq.ShortSeq192:
        return list((ctypes.c_uint64 * 3).from_address(id(obj) + 16))
    if type(obj) is sq.ShortSeqVar:
        ptr = ctypes.c_uint64.from_address(id(obj) + 16).value
        nb = (n + 31) // 32
        return list((ctypes.c_uint64 * nb).from_address(ptr))
    raise TypeError(type(obj))
