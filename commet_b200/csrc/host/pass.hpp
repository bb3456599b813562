// What index_and_search and compare_reads share on the host side: directory checks, loading the files of a set,
// and ONE chunk loop (src/index_and_search.cpp:255-277, src/compare_reads.cpp:248-259) through the C-ABI.
#pragma once
#include <sys/stat.h>
#include <sys/types.h>

#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "commet_b200.h"
#include "read_set.hpp"
#include "set_parser.hpp"

namespace commet_host {

inline void ensure_dir(const std::string &path)        // src/index_and_search.cpp:178-191
{
    struct stat info;
    if (stat(path.c_str(), &info) != 0) {
        mkdir(path.c_str(), S_IRWXU | S_IRGRP | S_IXGRP);
    } else if (!(info.st_mode & S_IFDIR)) {
        std::cerr << "Error: " << path << " already exists and is not a directory\n";
        exit(1);
    }
}

inline void load_set(ReadSet &set, const SetSpec &spec)
{
    for (size_t i = 0; i < spec.files.size(); i++) {
        if (spec.bvs[i].empty()) std::cout << "open " << spec.files[i] << "\n";
        else std::cout << "open " << spec.files[i] << "," << spec.bvs[i] << "\n";
        set.add_file(spec.files[i], spec.bvs[i]);
    }
}

struct PassResult {
    uint64_t indexed = 0, chunks = 0;
    std::vector<uint64_t> searched, shared;
    double index_s = 0, search_s = 0, total_s = 0;
};

// One chunk loop (src/index_and_search.cpp:255-277) on the GPU: `index` against every set of `queries`.
inline PassResult run_pass(commet_ctx *ctx, int k, int t, uint64_t max_kmer, ReadSet &index,
                           std::vector<ReadSet *> &queries, bool banners, const char *tool = "index_and_search")
{
    PassResult res;
    size_t ns = queries.size();
    res.searched.assign(ns, 0);
    res.shared.assign(ns, 0);
    std::vector<const uint8_t *> qb(ns);
    std::vector<const uint64_t *> qo(ns);
    std::vector<uint64_t> nq(ns);
    std::vector<std::vector<uint8_t>> tags(ns);
    std::vector<uint8_t *> tp(ns);
    static const uint8_t none = 0;
    for (size_t s = 0; s < ns; s++) {
        qb[s] = queries[s]->bases.empty() ? &none : queries[s]->bases.data();
        qo[s] = queries[s]->offs.data();
        nq[s] = queries[s]->n_valid();
        tags[s].assign(nq[s] / 8 + 1, 0);
        tp[s] = tags[s].data();
    }
    uint64_t stats[8] = {0};
    auto t0 = std::chrono::steady_clock::now();
    int rc = commet_index_and_search(ctx, k, t, max_kmer, index.bases.empty() ? &none : index.bases.data(),
                                     index.offs.data(), index.n_valid(), (int)ns, qb.data(), qo.data(), nq.data(),
                                     tp.data(), res.searched.data(), res.shared.data(), stats);
    if (rc != 0) {
        std::cerr << tool << ": " << commet_last_error() << "\n";
        exit(1);
    }
    res.total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    res.indexed = stats[1];
    res.chunks = stats[0];
    res.index_s = stats[3] * 1e-9;
    res.search_s = stats[4] * 1e-9;
    if (banners) {
        // the reference prints one banner per chunk and query set (:267-269)
        for (uint64_t ch = 0; ch < stats[0]; ch++)
            for (size_t s = 0; s < ns; s++) {
                std::cout << "\n------------------------------------------------------------------\n";
                std::cout << "finding reads from {" << queries[s]->nickname << "} present in raw {" << index.nickname << "}\n";
                std::cout << "------------------------------------------------------------------\n";
            }
    }
    for (size_t s = 0; s < ns; s++) queries[s]->scatter_tags(tags[s]);
    return res;
}


}  // namespace commet_host
