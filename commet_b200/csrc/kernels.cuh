// Device code of the B200-native Commet hot path (sm_100a).
//
// Data layout in HBM:
//   planes : uint4 per 32 consecutive bases of a read stream
//            .x = H  bit-plane (1 for G,T)   -> key a   (hash_key.h:65-91)
//            .y = L  bit-plane (1 for C,T)   -> key b ; c = H^L ; d = H|L
//            .z = V  validity  (1 for ACGTacgt, alphabet.h:44-58)
//            .w = W  "a k-mer starts here" for the k the stream was prepared
//                    for: k valid bases that do not cross a read boundary
//            bit j of a word = base 32*word + j (LSB first).
//   filter : the bloom_filter.h byte array viewed as little-endian u32 words:
//            key -> word key>>3, bit 8*((key>>1)&3) + (key&1 ? 3-j : 7-j).
//   tags   : u32 words, bit r%32 of word r/32 = read r (= .bv payload bytes).
//
// The kernels live in kernels/*.cuh, one file per stage of the path:
//   kernels/common.cuh  loads, the four keys of a k-mer as plane windows, their place in the filter (hash_key.h:65-125, bloom_filter.h:112-131)
//   kernels/staging.cuh  ASCII -> H/L/V bit-planes, read starts, the W plane, per-read k-mer counts (alphabet.h:44-58, index_reads.h:52-58)
//   kernels/insert.cuh  stage 1: the filter of a chunk -- direct RED.OR, the L2-blocked insert in its two forms, region passes (bloom_filter.h:112-121, index_reads.h:41-63)
//   kernels/search.cuh  stage 2: greedy non-overlapping k-mer search on both strands (search_reads.h:34-87, bloom_filter.h:124-131)
//   kernels/filter.cuh  stage 3: filter_reads -- length, N, Shannon, -m cutoff (filter_reads.cpp:181-205,265-306)
//   kernels/bvop.cuh  stage 4: boolean-vector operators and nb_one (boolean_vector.h:244-270,418-462)
//   kernels/multi.cuh  multi-GPU: merge of partial filters, owner-applied insert, region all-gather, over NVLink peer memory
//   kernels/ubench.cuh  measurement kernels: random 32-byte-sector loads / RED.OR (the ceilings bench.py reports against)
#pragma once
#include "kernels/common.cuh"
#include "kernels/staging.cuh"
#include "kernels/insert.cuh"
#include "kernels/search.cuh"
#include "kernels/filter.cuh"
#include "kernels/bvop.cuh"
#include "kernels/multi.cuh"
#include "kernels/ubench.cuh"
