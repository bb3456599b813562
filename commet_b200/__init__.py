"""commet_b200 -- B200-native implementation of Commet's index_and_search hot path.

Host-side mirror of the reference's interfaces (index_reads, search_reads,
filter_reads, bvop) over the C-ABI in include/commet_b200.h.  The CUDA
extension is mandatory: importing works anywhere, but creating a Context
without the built library or without a GPU raises -- there is no CPU fallback.
"""
from .api import (BV_AND, BV_ANDNOT, BV_NOT, BV_OR, CommetError, Context, Dist, Group, ReadStream, filter_bytes, lib_path,
                  load_library, max_kmer)

__all__ = ["Context", "Dist", "Group", "ReadStream", "CommetError", "filter_bytes", "max_kmer", "load_library", "lib_path",
           "BV_AND", "BV_OR", "BV_ANDNOT", "BV_NOT"]
