"""shortseq_b200 -- the B200-native ShortSeq hot path.

Same public names as the reference's `shortseq` package (shortseq/__init__.py:1-14):
pack / from_str / from_bytes, ShortSeq64 / ShortSeq192 / ShortSeqVar, ShortSeqCounter,
read_and_count_fastq, get_domain_* and the MIN/MAX_*_NT constants -- plus the batch entry
points that are the point of this package: pack_batch, decode_batch, hamming_batch,
hamming_refset, DeviceCounter, decode_many / hamming_many (str() and ^ over lists of boxed objects with one kernel
launch per class) and the multi-GPU merge in shortseq_b200.distributed.

All arithmetic runs in hand-written CUDA kernels for sm_100a behind a C ABI
(include/shortseq_b200.h, libshortseq_b200.so).  There is no CPU fallback.
"""
from ._lib import CLASS_64, CLASS_192, CLASS_VAR, LibraryError
from ._runtime import ShortSeqClassError
from .short_seq import (ShortSeq64, ShortSeq192, ShortSeqVar, pack, from_str, from_bytes, empty, decode_many, hamming_many,
                        get_domain_64, get_domain_192, get_domain_var)
from .batch import (ReadBatch, ShortSeqArray, pack_batch, pack_mixed, decode_batch, hamming_batch, hamming_refset,
                    synth_reads, slice_batch, kmers_batch, normalize_batch, umi_collapse)
from .counter import DeviceCounter, ShortSeqCounter, read_and_count_fastq

MIN_VAR_NT, MAX_VAR_NT = get_domain_var()
MIN_192_NT, MAX_192_NT = get_domain_192()
MIN_64_NT, MAX_64_NT = get_domain_64()

__all__ = [
    "pack", "from_str", "from_bytes", "ShortSeq64", "ShortSeq192", "ShortSeqVar", "ShortSeqCounter",
    "read_and_count_fastq", "get_domain_64", "get_domain_192", "get_domain_var",
    "MIN_64_NT", "MAX_64_NT", "MIN_192_NT", "MAX_192_NT", "MIN_VAR_NT", "MAX_VAR_NT",
    "pack_batch", "pack_mixed", "decode_batch", "hamming_batch", "hamming_refset", "synth_reads",
    "ReadBatch", "ShortSeqArray", "DeviceCounter", "CLASS_64", "CLASS_192", "CLASS_VAR",
    "LibraryError", "ShortSeqClassError", "empty", "decode_many", "hamming_many", "slice_batch", "kmers_batch", "normalize_batch", "umi_collapse",
]
