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
