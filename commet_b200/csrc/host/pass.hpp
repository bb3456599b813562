// What index_and_search and compare_reads share on the host side: directory checks, loading the files of a set,
// and ONE chunk loop (src/index_and_search.cpp:255-277, src/compare_reads.cpp:248-259) through the C-ABI.
#pragma once
#include <sys/stat.h>
#include <sys/types.h>

#include <chrono>
#include <cstdlib>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "commet_b200.h"
#include "read_set.hpp"
#include "set_parser.hpp"

namespace commet_host {

// Background threads of a tool (CUDA start-up, files parsed ahead) must not be running when exit() -- the reference's
// way out of every fatal error -- starts tearing the process down: they are joined by an atexit handler.
inline std::vector<std::thread *> &tool_threads()
{
    static std::vector<std::thread *> *v = new std::vector<std::thread *>;      // never destroyed: used by an atexit handler
    return *v;
}
inline void join_tool_threads()
{
    for (std::thread *t : tool_threads())
        if (t->joinable() && t->get_id() != std::this_thread::get_id()) t->join();
}
inline void watch_thread(std::thread *t)
{
    static bool registered = false;
    if (!registered) { atexit(join_tool_threads); registered = true; }
    tool_threads().push_back(t);
}
inline void unwatch_thread(std::thread *t)           // the owner has joined it and is going away
{
    std::vector<std::thread *> &v = tool_threads();
    v.erase(std::remove(v.begin(), v.end(), t), v.end());
}

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

// The search sets' large plain-FASTA files parsed by a background thread while the main thread loads the index set
// (the reference loads them one after the other, src/index_and_search.cpp:196-234).  Only the silent fast path runs
// ahead (fast_fasta.hpp); anything else -- other formats, unreadable files, a malformed fof -- is left to the main
// thread, which reports it where the reference would.
struct ParseAhead {
    std::thread th;
    std::map<std::string, ParsedFile> done;

    void start(const std::string &fof)
    {
        th = std::thread([this, fof]() {
            std::ifstream in(fof.c_str());
            std::string line;
            while (in.good() && std::getline(in, line)) {
                const size_t colon = line.find(':');
                if (colon == std::string::npos) continue;
                std::string rest = line.substr(colon + 1);
                size_t p = 0;
                while (p <= rest.size()) {
                    size_t q = rest.find(';', p);
                    if (q == std::string::npos) q = rest.size();
                    std::string item = rest.substr(p, q - p);
                    const size_t comma = item.find(',');
                    if (comma != std::string::npos) item = item.substr(0, comma);
                    const size_t a = item.find_first_not_of(" \t\r"), b = item.find_last_not_of(" \t\r");
                    if (a != std::string::npos) {
                        const std::string fname = item.substr(a, b - a + 1);
                        ParsedFile pf;
                        pf.fname = fname;
                        if (!done.count(fname) && parse_fasta_parallel(fname, pf)) done[fname] = std::move(pf);
                    }
                    p = q + 1;
                }
            }
        });
        watch_thread(&th);
    }
    ParsedFile *take(const std::string &fname)
    {
        if (th.joinable()) th.join();
        auto it = done.find(fname);
        return it == done.end() ? nullptr : &it->second;
    }
    ~ParseAhead() { if (th.joinable()) th.join(); unwatch_thread(&th); }
};

inline void load_set(ReadSet &set, const SetSpec &spec, ParseAhead *ahead = nullptr)
{
    for (size_t i = 0; i < spec.files.size(); i++) {
        if (spec.bvs[i].empty()) std::cout << "open " << spec.files[i] << "\n";
        else std::cout << "open " << spec.files[i] << "," << spec.bvs[i] << "\n";
        set.add_file(spec.files[i], spec.bvs[i], ahead ? ahead->take(spec.files[i]) : nullptr);
    }
}

struct PassResult {
    uint64_t indexed = 0, chunks = 0;
    std::vector<uint64_t> searched, shared;
    double index_s = 0, search_s = 0, total_s = 0;
};

// The GPU(s) a tool runs on.  COMMET_B200_GPUS = N asks for the first N visible devices, "all" for every one; unset:
// one GPU, or every visible one when the sets hold at least 2e9 bases (what makes dealing the index set over the
// box worth the extra contexts).  COMMET_B200_DEVICES = "0,0,1" names the devices rank by rank (tests: several ranks
// may share a GPU).  More than one device -> commet_group_* (the index set dealt over the GPUs, the partial filters
// merged over NVLink); otherwise a plain context on device 0.
struct Engine {
    commet_ctx *ctx = nullptr;
    commet_group *group = nullptr;
    std::thread warm;                       // creates the context of the first device while the read files are parsed
    commet_ctx *warm_ctx = nullptr;
    std::string warm_err;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();

    void mark(const char *what) const       // COMMET_B200_TRACE=1: wall clock of the tool's stages on stderr
    {
        if (!getenv("COMMET_B200_TRACE")) return;
        std::cerr << "[commet tool] +" << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()
                  << " ms " << what << "\n";
    }
    static std::vector<int> named_devices()
    {
        std::vector<int> devices;
        if (const char *e = getenv("COMMET_B200_DEVICES"))
            for (const char *p = e; *p;) {
                devices.push_back(atoi(p));
                while (*p && *p != ',') p++;
                if (*p == ',') p++;
            }
        return devices;
    }
    // CUDA initialisation and the context of the first device take a few hundred milliseconds: started before the
    // files are read, joined by open()
    void prewarm()
    {
        std::vector<int> named = named_devices();
        const int dev = named.empty() ? 0 : named[0];
        warm = std::thread([this, dev]() {
            if (commet_ctx_create(dev, &warm_ctx) != 0) warm_err = commet_last_error();
        });
        watch_thread(&warm);
    }
    bool open(uint64_t total_bases)
    {
        if (warm.joinable()) warm.join();
        mark("context of the first device ready");
        std::vector<int> devices = named_devices();
        if (devices.empty()) {
            int want = 1;
            const int visible = commet_device_count();
            const char *e2 = getenv("COMMET_B200_GPUS");
            if (e2 && std::string(e2) == "all") want = visible;
            else if (e2 && atoi(e2) > 0) want = atoi(e2);
            else if (total_bases >= 2000000000ull) want = visible;
            want = std::max(1, std::min(std::min(want, visible), 8));
            for (int i = 0; i < want; i++) devices.push_back(i);
        }
        if (devices.size() > 1) {
            if (warm_ctx) { commet_ctx_destroy(warm_ctx); warm_ctx = nullptr; }      // the group creates its own contexts
            return commet_group_create(devices.data(), (int)devices.size(), &group) == 0;
        }
        if (warm_ctx) { ctx = warm_ctx; warm_ctx = nullptr; return true; }
        if (!warm_err.empty()) return commet_ctx_create(devices.empty() ? 0 : devices[0], &ctx) == 0;   // repeats the failure: sets the message in this thread
        return commet_ctx_create(devices.empty() ? 0 : devices[0], &ctx) == 0;
    }
    ~Engine() { if (warm.joinable()) warm.join(); unwatch_thread(&warm); }
    void close()
    {
        if (warm.joinable()) warm.join();
        if (warm_ctx) commet_ctx_destroy(warm_ctx);
        if (group) commet_group_destroy(group);
        if (ctx) commet_ctx_destroy(ctx);
        group = nullptr;
        ctx = nullptr;
        warm_ctx = nullptr;
    }
};

// One chunk loop (src/index_and_search.cpp:255-277) on the GPU(s): `index` against every set of `queries`.
inline PassResult run_pass(Engine &eng, int k, int t, uint64_t max_kmer, ReadSet &index,
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
    const uint8_t *ib = index.bases.empty() ? &none : index.bases.data();
    int rc = eng.group ? commet_group_index_and_search(eng.group, k, t, max_kmer, ib, index.offs.data(), index.n_valid(), (int)ns,
                                                       qb.data(), qo.data(), nq.data(), tp.data(), res.searched.data(),
                                                       res.shared.data(), stats)
                       : commet_index_and_search(eng.ctx, k, t, max_kmer, ib, index.offs.data(), index.n_valid(), (int)ns,
                                                 qb.data(), qo.data(), nq.data(), tp.data(), res.searched.data(),
                                                 res.shared.data(), stats);
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
