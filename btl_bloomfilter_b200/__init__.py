"""btl_bloomfilter_b200 -- the k-mer insert/query hot path of bcgsc/btl_bloomfilter on B200 (sm_100a).

Host-side mirror of the reference's class API (BloomFilter, CountingBloomFilter<uint8_t>, insertSeq)
plus batched insertSeqs / containsSeqs, over the C ABI of include/btlbf.h (libbtlbf_cuda.so: hand-written
CUDA kernels).  Importing the classes requires the built CUDA library; there is no CPU fallback.
"""
from ._build import build_library  # noqa: F401
from ._capi import BLOOM, COUNTING8, BtlbfError, lib  # noqa: F401
from .filters import (BitVector, BloomFilter, Context, CountingBloomFilter, KmerBloomFilter, PackedBatch,  # noqa: F401
                      QueryResult, as_batch, insertSeq, pack_seqs, unpack_bits)

__all__ = ["BitVector", "BloomFilter", "CountingBloomFilter", "KmerBloomFilter", "Context", "QueryResult", "insertSeq", "as_batch",
           "unpack_bits", "PackedBatch", "pack_seqs", "build_library", "lib", "BtlbfError", "BLOOM", "COUNTING8"]
