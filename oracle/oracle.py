"""ctypes binding of oracle/libssq_oracle.so (the C restatement in ssq_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by shortseq_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, ERR_BAD_BASE, ERR_TOO_LONG, ERR_CLASS, ERR_LEN_MISMATCH, ERR_UB = 0, 1, 2, 3, 5, 100


class Err(C.Structure):
    _fields_ = [("code", C.c_int32), ("bad_len", C.c_int32), ("bad_pos", C.c_int64),
                ("bad_chars", C.c_uint8 * 8)]

    def chars(self):
        return bytes(self.bad_chars[: self.bad_len])


def build():
    so = os.path.join(HERE, "libssq_oracle.so")
    src = os.path.join(HERE, "ssq_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "libssq_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        p = C.c_void_p
        L.ssq_oracle_pack_one.argtypes = [p, C.c_size_t, p, C.POINTER(Err)]
        L.ssq_oracle_pack_batch.argtypes = [C.c_int, p, p, C.c_int64, p, p, p,
                                            C.POINTER(C.c_int64), C.POINTER(Err)]
        L.ssq_oracle_hamming_batch.argtypes = [p, p, p, p, C.c_int64, C.c_size_t, p,
                                               C.POINTER(C.c_int64)]
        L.ssq_oracle_decode_batch.argtypes = [p, p, C.c_size_t, p, C.c_int64, p, p]
        L.ssq_oracle_decode_batch.restype = None
        L.ssq_oracle_count.argtypes = [p, p, C.c_int64, C.c_size_t, p, p, p, p]
        L.ssq_oracle_count.restype = C.c_int64
        L.ssq_oracle_pyhash.argtypes = [C.c_uint64]
        L.ssq_oracle_pyhash.restype = C.c_int64
        L.ssq_oracle_is_base.argtypes = [C.c_uint8]
        L.ssq_oracle_container_words.argtypes = [C.c_size_t]
        L.ssq_oracle_container_words.restype = C.c_size_t
        L.ssq_oracle_class.argtypes = [C.c_size_t]
        L.ssq_oracle_mix64.argtypes = [C.c_uint64]
        L.ssq_oracle_mix64.restype = C.c_uint64
        L.ssq_oracle_synth_reads.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64,
                                             C.c_int32, C.c_int32, p, p]
        L.ssq_oracle_synth_reads.restype = C.c_int64
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(Exception):
    def __init__(self, code, first_bad, err):
        self.code, self.first_bad, self.bad_chars, self.bad_pos = code, first_bad, err.chars(), err.bad_pos
        super().__init__(f"oracle status {code} at read {first_bad}: {self.bad_chars!r}")


def concat(reads):
    """list of bytes -> (uint8 buffer, int64 offsets[n+1])."""
    lens = np.fromiter((len(r) for r in reads), dtype=np.int64, count=len(reads))
    offsets = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    buf = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if len(reads) else np.zeros(0, np.uint8)
    return buf, offsets


def pack_one(seq: bytes):
    """-> (klass, words[np.uint64]) or raises OracleError."""
    L = lib()
    nw = max(1, L.ssq_oracle_container_words(len(seq)))
    words = np.zeros(nw, dtype=np.uint64)
    err = Err()
    buf = np.frombuffer(seq, dtype=np.uint8) if seq else np.zeros(1, np.uint8)
    rc = L.ssq_oracle_pack_one(_ptr(buf), len(seq), _ptr(words), C.byref(err))
    if rc:
        raise OracleError(rc, 0, err)
    return L.ssq_oracle_class(len(seq)), words


def pack_batch(klass, ascii_buf, offsets):
    """Class-homogeneous batch -> (words, lens[int32], word_off or None)."""
    L = lib()
    n = len(offsets) - 1
    lens_in = np.diff(offsets)
    if klass == 0:
        words = np.zeros(n, np.uint64)
    elif klass == 1:
        words = np.zeros((n, 3), np.uint64)
    else:
        words = np.zeros(int(((np.clip(lens_in, 0, 1024) + 31) // 32).sum()) + 1, np.uint64)
    lens = np.zeros(n, np.int32)
    word_off = np.zeros(n + 1, np.int64) if klass == 2 else None
    first_bad = C.c_int64(-1)
    err = Err()
    a = ascii_buf if len(ascii_buf) else np.zeros(1, np.uint8)
    rc = L.ssq_oracle_pack_batch(klass, _ptr(a), _ptr(offsets), n, _ptr(words), _ptr(lens),
                                 _ptr(word_off), C.byref(first_bad), C.byref(err))
    if rc:
        raise OracleError(rc, first_bad.value, err)
    if klass == 2:
        words = words[: word_off[n]]
    return words, lens, word_off


def hamming_batch(a, b, len_a, len_b, stride):
    L = lib()
    n = len(len_a)
    dist = np.zeros(n, np.int32)
    fb = C.c_int64(-1)
    rc = L.ssq_oracle_hamming_batch(_ptr(np.ascontiguousarray(a)), _ptr(np.ascontiguousarray(b)),
                                    _ptr(len_a.astype(np.int32)), _ptr(len_b.astype(np.int32)),
                                    n, stride, _ptr(dist), C.byref(fb))
    if rc:
        raise OracleError(rc, fb.value, Err())
    return dist


def decode_batch(words, lens, stride=1, word_off=None):
    """-> (uint8 ascii, int64 out_off[n+1])."""
    L = lib()
    lens = lens.astype(np.int32)
    n = len(lens)
    out_off = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=out_off[1:])
    out = np.zeros(max(1, int(out_off[n])), np.uint8)
    L.ssq_oracle_decode_batch(_ptr(np.ascontiguousarray(words)), _ptr(word_off), stride, _ptr(lens), n,
                              _ptr(out), _ptr(out_off))
    return out[: out_off[n]], out_off


def count(words, lens, stride):
    """-> (uniq_words, uniq_lens, counts, first_idx) in first-occurrence order."""
    L = lib()
    n = len(lens)
    words = np.ascontiguousarray(words, dtype=np.uint64)
    lens = lens.astype(np.int32)
    uw = np.zeros((max(n, 1), stride), np.uint64)
    ul = np.zeros(max(n, 1), np.int32)
    cnt = np.zeros(max(n, 1), np.int64)
    fi = np.zeros(max(n, 1), np.int64)
    u = L.ssq_oracle_count(_ptr(words), _ptr(lens), n, stride, _ptr(uw), _ptr(ul), _ptr(cnt), _ptr(fi))
    assert u >= 0
    uw = uw[:u]
    return (uw[:, 0] if stride == 1 else uw), ul[:u], cnt[:u], fi[:u]


def pyhash(word0):
    return lib().ssq_oracle_pyhash(int(word0))


def synth_reads(seed, first_read, n, n_keys, len_lo, len_hi):
    """Counter-based synthetic batch -> (ascii uint8, offsets int64[n+1])."""
    L = lib()
    offsets = np.zeros(n + 1, np.int64)
    total = L.ssq_oracle_synth_reads(seed, first_read, n, n_keys, len_lo, len_hi, None, _ptr(offsets))
    buf = np.zeros(max(1, total), np.uint8)
    L.ssq_oracle_synth_reads(seed, first_read, n, n_keys, len_lo, len_hi, _ptr(buf), _ptr(offsets))
    return buf[:total], offsets
