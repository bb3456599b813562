// C-ABI of the B200-native Commet hot path: context, read staging, chunk
// loop, and the launchers of the kernels in kernels/*.cuh -- one translation
// unit, its parts in capi/*.inl (included below, in order).  See
// include/commet_b200.h for the contract of every entry point and the
// reference interface it replaces.  There is no CPU fallback in this library:
// every data-path operation is a kernel launch on the context's stream.
#include "../../include/commet_b200.h"
#include "kernels.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace commet;

#include "capi/core.inl"          // errors, trace, the context's device-memory arena, the context and read-stream types, context life cycle
#include "capi/reads.inl"          // read staging (H2D copies, encode, W plane), read selection, the chunk plan of a staged set
#include "capi/index.inl"          // stage 1: the insert of a range of reads into the filter (direct, L2-blocked), filter transfer, merge of partial filters
#include "capi/search.inl"          // stage 2: the search launches and the chunk loop (commet_index_and_search and its staged / resident forms)
#include "capi/filter.inl"          // stage 3: filter_reads (selection on planes, fused staging + selection, host decisions for undecided reads)
#include "capi/bvop.inl"          // stage 4: boolean-vector operators, popcounts; the measurement entry point

#include "capi/dist.inl"          // the chunk loop over several GPUs (commet_dist_*, commet_group_*)
