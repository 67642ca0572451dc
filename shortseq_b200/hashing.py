"""Host-side (numpy) statement of the counter's key hash: which hash partition / rank owns a key.

The device tables mix their own slot hash (the reference's dict hash, word0, is never
observable): a fold-and-multiply bijection of the packed word for ShortSeq64 keys, a rotate-fold of the three
words and the length followed by one splitmix64 round for ShortSeq192 keys (csrc/ssq_device.cuh).  The owner of a key in a
P-way multi-GPU merge is the top log2(P) bits of that hash.  This module only computes
partition ids for bookkeeping and tests; it counts nothing.
"""
import numpy as np

_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def mix64(x):
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= _C1
        x ^= x >> np.uint64(27)
        x *= _C2
        x ^= x >> np.uint64(31)
    return x


def key_hash(words, lens, klass):
    """64-bit key hash per read; words [n] (ShortSeq64) or [n, 3] (ShortSeq192)."""
    words = np.asarray(words, dtype=np.uint64)
    if klass == 0:
        with np.errstate(over="ignore"):
            return (words ^ (words >> np.uint64(32))) * _GOLD        # hash64 of csrc/ssq_device.cuh
    def rotl(x, r):
        return (x << np.uint64(r)) | (x >> np.uint64(64 - r))
    with np.errstate(over="ignore"):
        return mix64(words[:, 0] ^ rotl(words[:, 1], 21) ^ (rotl(words[:, 2], 43) * _GOLD)
                     ^ (np.asarray(lens, dtype=np.uint64) << np.uint64(56)))


def owner_rank(words, lens, klass, world):
    """Rank that owns each key in a `world`-way merge (world a power of two)."""
    if world & (world - 1):
        raise ValueError("world size must be a power of two")
    if world == 1:
        return np.zeros(len(np.asarray(lens)), dtype=np.int64)
    bits = world.bit_length() - 1
    return (key_hash(words, lens, klass) >> np.uint64(64 - bits)).astype(np.int64)


# ---- the ShortSeq64 table key and its journey through the streamed multi-GPU exchange (csrc/ssq_table.cuh, ssq_counter.cu) ----
_MASK58 = np.uint64((1 << 58) - 1)


def _rotl(x, r):
    x = np.asarray(x, dtype=np.uint64)
    r %= 64
    return x if r == 0 else (x << np.uint64(r)) | (x >> np.uint64(64 - r))


def table_key64(words, lens, rot):
    """(h2, key) of ShortSeq64 keys in a table with hash rotation `rot`: h2 = rotl(hash64(word), rot); the stored key is
    h2's low 58 bits with len + 1 in bits 58..63 (key64_of); the home slot of a 2^c-slot table is h2 >> (64 - c)."""
    h2 = _rotl(key_hash(words, lens, 0), rot)
    key = (h2 & _MASK58) | ((np.asarray(lens, dtype=np.uint64) + np.uint64(1)) << np.uint64(58))
    return h2, key


def streamed_key64(local_key, region, log2_regions, rot_local, rot_owner):
    """What count_regions2_kernel<., true> sends for a slot of table region `region` holding `local_key`: the key in the
    OWNER table's format.  The six hash bits a key does not store are the top six bits of its region index."""
    local_key = np.asarray(local_key, dtype=np.uint64)
    top6 = (np.asarray(region, dtype=np.uint64) >> np.uint64(log2_regions - 6)) << np.uint64(58)
    h = _rotl(top6 | (local_key & _MASK58), 64 - rot_local)          # hash64(word)
    return (_rotl(h, rot_owner) & _MASK58) | (local_key & ~_MASK58)
