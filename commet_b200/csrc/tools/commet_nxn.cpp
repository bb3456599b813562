// commet_nxn -- the whole non-SGE run of Commet.py (Commet.py:438-598) in ONE process on resident read sets:
//   filtering            Commet.py:103-121   one filter_reads per file
//   N x N comparison     Commet.py:186-240   for every reference set i: "all in Si", then for every j > i
//                                            "Si in (Sj in Si)" and "Sj in (Si in (Sj in Si))": N^2-1 index_and_search
//   matrices             Commet.py:245-317   matrix_plain / matrix_percentage / matrix_normalized .csv
// Every read file is parsed once, every set is staged once per GPU (all records, 2-bit planes), and each
// index_and_search round is one commet_index_and_search_resident call under the round's boolean vectors
// (commet_reads_select) -- no process start, no re-parse, no re-upload, no `bvop -i` fork per matrix cell.
// Outputs are the reference's: <out>/<file>.bv (filter), <out>/<file>_in_<set>.bv, the <Q>_in_<I>.log counters and
// the three CSVs, byte-identical to what the unchanged Commet.py writes with the reference binaries (the CSV
// numbers are formatted like Python 3's str(float)).
// Rounds are independent once their input vectors exist, so they are spread over the visible GPUs: one worker
// thread and one context per device, "all in Si" rounds first, each (i, j) pair's two refinement rounds as soon
// as "all in Si" is done.  There is no CPU path: all selection, indexing, search and counting runs on the GPUs.
#include <sys/stat.h>
#include <sys/types.h>

#include <atomic>
#include <charconv>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "bv.hpp"
#include "commet_b200.h"
#include "fast_fasta.hpp"
#include "readers.hpp"

using namespace commet_host;

// ---------------------------------------------------------------- small helpers ----
static std::string strip(const std::string &s)          // Python str.strip()
{
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b])) b++;
    while (e > b && isspace((unsigned char)s[e - 1])) e--;
    return s.substr(b, e - b);
}

static std::string basename_of(const std::string &p)    // os.path.basename
{
    size_t pos = p.rfind('/');
    return pos == std::string::npos ? p : p.substr(pos + 1);
}

static std::vector<std::string> split(const std::string &s, char sep)
{
    std::vector<std::string> out;
    size_t b = 0;
    while (true) {
        size_t e = s.find(sep, b);
        if (e == std::string::npos) { out.push_back(s.substr(b)); break; }
        out.push_back(s.substr(b, e - b));
        b = e + 1;
    }
    return out;
}

// Python 3 str(float) (= repr): shortest digits that round-trip; exponent form when the decimal point position
// is <= -4 or > 16 (Python/pystrtod.c, format_float_short, 'r'), at least two exponent digits, ".0" otherwise
static std::string py_float(double x)
{
    if (x == 0) return std::signbit(x) ? "-0.0" : "0.0";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
    std::string sci(buf, res.ptr);                      // [-]d[.ddd]e[+-]XX
    std::string sign;
    if (sci[0] == '-') { sign = "-"; sci = sci.substr(1); }
    size_t e = sci.find('e');
    std::string mant = sci.substr(0, e);
    int exp10 = atoi(sci.c_str() + e + 1);
    std::string digits;
    for (char ch : mant) if (ch != '.') digits += ch;
    int decpt = exp10 + 1;                              // value = 0.digits x 10^decpt
    std::string out = sign;
    if (decpt <= -4 || decpt > 16) {
        out += digits[0];
        if (digits.size() > 1) out += "." + digits.substr(1);
        int ex = decpt - 1;
        out += ex < 0 ? "e-" : "e+";
        ex = ex < 0 ? -ex : ex;
        if (ex < 10) out += "0";
        out += std::to_string(ex);
    } else if (decpt <= 0) {
        out += "0." + std::string((size_t)(-decpt), '0') + digits;
    } else if ((size_t)decpt >= digits.size()) {
        out += digits + std::string((size_t)decpt - digits.size(), '0') + ".0";
    } else {
        out += digits.substr(0, (size_t)decpt) + "." + digits.substr((size_t)decpt);
    }
    return out;
}

static void ensure_dir(const std::string &path)
{
    struct stat info;
    if (stat(path.c_str(), &info) != 0) mkdir(path.c_str(), 0755);
}

[[noreturn]] static void die(const std::string &msg)
{
    std::cerr << "commet_nxn: " << msg << "\n";
    exit(1);
}

// ------------------------------------------------------------------- data model ----
struct SetFileInfo {
    std::string path;                 // as written in the config (the .bv comments embed it)
    std::string bv_path;              // filter vector: given in the config or <out>/<basename>.bv
    uint64_t first = 0, n = 0;        // records [first, first+n) of the set stream
};

struct ReadSetData {
    std::string name;
    std::vector<SetFileInfo> files;
    uint8_t *bases = nullptr;         // all records of all files, concatenated (pinned when possible)
    bool pinned = false;
    uint64_t n_bases = 0;
    std::vector<uint64_t> offs{0};
    uint64_t n_records() const { return offs.size() - 1; }
};

// the boolean vectors the reference flow keeps as files, keyed by the path Commet.py would use
struct VectorStore {
    std::mutex mu;
    std::map<std::string, BitVec> v;
    BitVec get(const std::string &path)
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = v.find(path);
        if (it == v.end()) die("boolean vector " + path + " is not available");
        return it->second;
    }
    void put(const std::string &path, const BitVec &bv)
    {
        std::lock_guard<std::mutex> g(mu);
        v[path] = bv;
    }
};

// concatenation of the files' vectors over the set's records (.bv payload layout)
static std::vector<uint8_t> concat_bits(const ReadSetData &set, const std::vector<BitVec> &per_file)
{
    std::vector<uint8_t> out(set.n_records() / 8 + 1, 0);
    for (size_t f = 0; f < set.files.size(); f++) {
        const SetFileInfo &sf = set.files[f];
        const BitVec &b = per_file[f];
        if (b.n != sf.n) {       // fasta_file.h:104-107
            std::cerr << "Number of reads in " << sf.path << " and boolean vector size are not equal -> quit\n";
            exit(1);
        }
        if ((sf.first & 7) == 0) {
            for (uint64_t i = 0; i < sf.n / 8; i++) out[sf.first / 8 + i] = b.bytes[i];
            for (uint64_t r = sf.n / 8 * 8; r < sf.n; r++)
                if (b.get(r)) out[(sf.first + r) / 8] |= (uint8_t)(1u << ((sf.first + r) % 8));
        } else {
            for (uint64_t r = 0; r < sf.n; r++)
                if (b.get(r)) out[(sf.first + r) / 8] |= (uint8_t)(1u << ((sf.first + r) % 8));
        }
    }
    return out;
}

static BitVec slice_bits(const std::vector<uint8_t> &bits, uint64_t first, uint64_t n)
{
    BitVec b;
    b.init_false(n);
    if ((first & 7) == 0) {
        for (uint64_t i = 0; i < n / 8; i++) b.bytes[i] = bits[first / 8 + i];
        for (uint64_t r = n / 8 * 8; r < n; r++)
            if ((bits[(first + r) / 8] >> ((first + r) % 8)) & 1u) b.set(r);
    } else {
        for (uint64_t r = 0; r < n; r++)
            if ((bits[(first + r) / 8] >> ((first + r) % 8)) & 1u) b.set(r);
    }
    return b;
}

// --------------------------------------------------------------------- options ----
struct Options {
    std::string config, out_dir = "output_commet/";
    int k = 33, t = 2, l = 0, n = -1, m = -1, gpus = 0;
    double e = 0;
    std::string e_text = "0";          // str(args.e) as Commet.py passes it on
    bool quiet = false;
    std::string report;                // --report <file>: phase timings as JSON
};

static void usage()
{
    std::cerr << "Usage: commet_nxn <sets_config> [-o dir] [-k 33] [-t 2] [-l 0] [-n any] [-e 0] [-m all] [--gpus N] [-q]\n"
                 "  same config format, options and outputs as `python Commet.py <sets_config> ...` (non-SGE mode)\n";
}

// one index_and_search invocation of the reference flow
struct Round {
    int index_set = -1;
    std::vector<std::string> index_bvs;            // one per file of the index set
    std::vector<int> query_sets;
    std::vector<std::vector<std::string>> query_bvs;
    bool write_logs = true;
};

struct RoundOut {
    uint64_t indexed = 0;
    std::vector<uint64_t> searched, shared, ones;
    double index_s = 0, search_s = 0, total_s = 0;
    uint64_t chunks = 0;
};

struct Driver {
    Options opt;
    std::vector<ReadSetData> sets;
    VectorStore store;
    uint64_t max_kmer = 0;
    std::mutex io_mu;

    struct Worker {
        int device = 0;
        commet_ctx *ctx = nullptr;
        std::vector<commet_reads *> staged;        // per set, lazily
    };
    std::vector<Worker> workers;

    // A set crosses PCIe once: the first GPU that needs it uploads and encodes it, every other GPU clones the
    // encoded planes from that GPU over NVLink (commet_reads_clone).
    std::mutex stage_mu;
    std::condition_variable stage_cv;
    std::vector<int> owner;                        // per set: -1 nobody, -2 being uploaded, g = staged on GPU g
    commet_reads *stage(Worker &w, int s)
    {
        if (w.staged[s]) return w.staged[s];
        ReadSetData &set = sets[s];
        int from = -1;
        {
            std::unique_lock<std::mutex> lk(stage_mu);
            if (owner.empty()) owner.assign(sets.size(), -1);
            stage_cv.wait(lk, [&]() { return owner[s] != -2; });
            from = owner[s];
            if (from == -1) owner[s] = -2;
        }
        auto t0 = std::chrono::steady_clock::now();
        if (from >= 0) {
            if (commet_reads_clone(w.ctx, workers[from].staged[s], &w.staged[s]) != 0)
                die(std::string("cloning set ") + set.name + ": " + commet_last_error());
            say("  set " + set.name + " cloned GPU " + std::to_string(from) + " -> GPU " + std::to_string(w.device) + " in " +
                std::to_string(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()) + " ms");
        } else {
            static const uint8_t none = 0;
            if (commet_reads_upload(w.ctx, set.n_bases ? set.bases : &none, set.offs.data(), set.n_records(), &w.staged[s]) != 0)
                die(std::string("staging set ") + set.name + ": " + commet_last_error());
            {
                std::lock_guard<std::mutex> lk(stage_mu);
                owner[s] = w.device;
            }
            stage_cv.notify_all();
            say("  set " + set.name + " uploaded to GPU " + std::to_string(w.device) + " in " +
                std::to_string(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()) + " ms");
        }
        return w.staged[s];
    }

    std::string out_bv_path(const SetFileInfo &f, const std::string &index_name) const
    {
        return opt.out_dir + basename_of(f.path) + "_in_" + index_name + ".bv";     // file_manager.h:247
    }

    void select(Worker &w, int s, const std::vector<std::string> &bv_paths)
    {
        std::vector<BitVec> per_file;
        for (const std::string &p : bv_paths) per_file.push_back(store.get(p));
        std::vector<uint8_t> bits = concat_bits(sets[s], per_file);
        if (commet_reads_select(w.ctx, stage(w, s), bits.data()) != 0) die(commet_last_error());
    }

    RoundOut run_round(Worker &w, const Round &r)
    {
        const size_t nq = r.query_sets.size();
        RoundOut out;
        out.searched.assign(nq, 0);
        out.shared.assign(nq, 0);
        out.ones.assign(nq, 0);
        select(w, r.index_set, r.index_bvs);
        std::vector<commet_reads *> q(nq);
        std::vector<std::vector<uint8_t>> tags(nq);
        std::vector<uint8_t *> tp(nq);
        for (size_t s = 0; s < nq; s++) {
            select(w, r.query_sets[s], r.query_bvs[s]);
            q[s] = stage(w, r.query_sets[s]);
            tags[s].assign(sets[r.query_sets[s]].n_records() / 8 + 1, 0);
            tp[s] = tags[s].data();
        }
        uint64_t stats[8] = {0};
        auto t0 = std::chrono::steady_clock::now();
        if (commet_index_and_search_resident(w.ctx, opt.k, opt.t, max_kmer, stage(w, r.index_set), (int)nq, q.data(), tp.data(),
                                             out.searched.data(), out.shared.data(), out.ones.data(), stats) != 0)
            die(commet_last_error());
        out.total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out.chunks = stats[0];
        out.indexed = stats[1];
        out.index_s = stats[3] * 1e-9;
        out.search_s = stats[4] * 1e-9;
        const std::string &iname = sets[r.index_set].name;
        for (size_t s = 0; s < nq; s++) {
            const ReadSetData &qs = sets[r.query_sets[s]];
            if (out.ones[s] != out.shared[s]) die("internal: tag popcount differs from the shared counter");
            for (const SetFileInfo &f : qs.files) {               // FileManager::save_bv, file_manager.h:245-252
                BitVec b = slice_bits(tags[s], f.first, f.n);
                b.comment = f.path + " in " + iname;
                store.put(out_bv_path(f, iname), b);
            }
            if (r.write_logs) {                                   // src/index_and_search.cpp:286-300
                std::ofstream log((opt.out_dir + qs.name + "_in_" + iname + ".log").c_str());
                log << "Index  time: " << (float)out.index_s << " s\n";
                log << "Search time: " << (float)out.search_s << " s\n";
                log << "Total  time: " << (float)out.total_s << " s\n";
                log << "[indexed " << out.indexed << ", searched " << out.searched[s] << ", shared " << out.shared[s] << "]\n";
            }
        }
        return out;
    }

    void say(const std::string &msg)
    {
        if (opt.quiet) return;
        std::lock_guard<std::mutex> g(io_mu);
        std::cout << msg << std::endl;
    }
};

// --------------------------------------------------------------------------- main ----
int main(int argc, char **argv)
{
    Driver d;
    Options &opt = d.opt;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) { usage(); exit(1); }
            return argv[++i];
        };
        if (a == "-o" || a == "--output_directory") opt.out_dir = val();
        else if (a == "-k") opt.k = atoi(val().c_str());
        else if (a == "-t") opt.t = atoi(val().c_str());
        else if (a == "-l") opt.l = atoi(val().c_str());
        else if (a == "-n") opt.n = atoi(val().c_str());
        else if (a == "-m") opt.m = atoi(val().c_str());
        else if (a == "-e") { opt.e = atof(val().c_str()); opt.e_text = py_float(opt.e); }
        else if (a == "--gpus") opt.gpus = atoi(val().c_str());
        else if (a == "-b" || a == "--binaries_directory") val();        // accepted for command-line compatibility
        else if (a == "-q") opt.quiet = true;
        else if (a == "--report") opt.report = val();
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (!a.empty() && a[0] == '-') { std::cerr << "Unknown option " << a << "\n"; usage(); return 1; }
        else if (opt.config.empty()) opt.config = a;
        else { usage(); return 1; }
    }
    if (opt.config.empty()) { usage(); return 1; }
    if (opt.out_dir.empty() || opt.out_dir.back() != '/') opt.out_dir += "/";
    if (opt.l < opt.k * opt.t && opt.l != 0) opt.l = opt.k * opt.t;       // Commet.py:509-513
    d.max_kmer = commet_max_kmer(opt.k);
    ensure_dir(opt.out_dir);
    auto t_start = std::chrono::steady_clock::now();
    std::vector<std::pair<std::string, double>> phases;       // --report: seconds since start at the end of each phase
    auto phase_done = [&](const char *name) {
        phases.emplace_back(name, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
    };

    // ---- config, Commet.py:42-95 ---------------------------------------------------------------------------
    std::ifstream cfg(opt.config.c_str());
    if (!cfg.good()) die("cannot read " + opt.config);
    bool with_bv = false;
    {
        std::string ln;
        bool first = true;
        std::vector<std::string> raw;
        while (std::getline(cfg, ln)) {
            if (first) { with_bv = ln.find(',') != std::string::npos; first = false; }    // getReadBVFiles looks at line 1 only
            if (strip(ln).empty()) continue;
            raw.push_back(ln);
        }
        for (const std::string &line : raw) {
            std::vector<std::string> parts = split(line, ':');
            if (parts.size() < 2) die("config line without ':' : " + line);
            ReadSetData set;
            set.name = strip(parts[0]);
            for (const std::string &item : split(parts[1], ';')) {
                std::vector<std::string> fb = split(strip(item), ',');
                SetFileInfo f;
                f.path = fb[0];
                if (with_bv) {
                    if (fb.size() < 2) die("config line without a boolean vector: " + line);
                    f.bv_path = fb[1];
                } else {
                    f.bv_path = opt.out_dir + basename_of(f.path) + ".bv";
                }
                set.files.push_back(f);
            }
            d.sets.push_back(std::move(set));
        }
        if (d.sets.empty()) die("no read set in " + opt.config);
    }
    // ---- one context per GPU, created in the background (0.6 s each) ------------------------------------------
    const bool parse_only = getenv("COMMET_NXN_PARSE_ONLY") != nullptr;     // host-side profiling of the readers
    if (!parse_only) {
        int n_dev = commet_device_count();
        if (n_dev <= 0) die("no CUDA device visible (there is no CPU path)");
        d.workers.resize(opt.gpus > 0 ? std::min(opt.gpus, n_dev) : n_dev);
    }
    // (one after the other: creating them from parallel threads was measured anywhere between 2x faster and 3x slower)
    std::thread ctx_thread([&]() {
        for (size_t g = 0; g < d.workers.size(); g++) {
            d.workers[g].device = (int)g;
            if (commet_ctx_create((int)g, &d.workers[g].ctx) != 0) die(commet_last_error());
        }
    });
    // ---- load every file once, straight into the set streams -------------------------------------------------
    // plain FASTA: mmap + chunked two-pass loader on all cores (fast_fasta.hpp); FASTQ / gzip: the sequential readers
    {
        struct Src { int s, f; bool fast = false; FastaMap map; ParsedFile slow; uint64_t base0 = 0; };
        std::vector<Src> src;
        for (size_t s = 0; s < d.sets.size(); s++)
            for (size_t f = 0; f < d.sets[s].files.size(); f++) {
                Src x;
                x.s = (int)s;
                x.f = (int)f;
                src.push_back(std::move(x));
            }
        const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
        auto parallel_for = [&](size_t n, const std::function<void(size_t)> &fn) {
            std::vector<std::thread> th;
            std::atomic<size_t> next{0};
            for (unsigned w = 0; w < std::min<size_t>(nt, n); w++)
                th.emplace_back([&]() {
                    for (size_t i; (i = next.fetch_add(1)) < n;) fn(i);
                });
            for (auto &t : th) t.join();
        };
        struct Item { size_t src, chunk; };
        std::vector<Item> items;
        for (size_t i = 0; i < src.size(); i++) {
            const std::string &path = d.sets[src[i].s].files[src[i].f].path;
            src[i].fast = src[i].map.open(path, 16u << 20);
            if (src[i].fast) for (size_t c = 0; c < src[i].map.n_chunks(); c++) items.push_back({i, c});
            else items.push_back({i, 0});
        }
        parallel_for(items.size(), [&](size_t it) {
            Src &x = src[items[it].src];
            if (x.fast) x.map.pass<false>(items[it].chunk, nullptr, 0, nullptr, 0);
            else if (!parse_reads_file(d.sets[x.s].files[x.f].path, x.slow, " -> quit\n")) exit(1);
        });
        phase_done("scan");
        size_t i = 0;
        for (ReadSetData &set : d.sets) {
            uint64_t total = 0, recs = 0;
            for (size_t f = 0; f < set.files.size(); f++, i++) {
                Src &x = src[i];
                if (x.fast) x.map.finish_scan();
                set.files[f].first = recs;
                set.files[f].n = x.fast ? x.map.n_records : x.slow.nb_reads;
                x.base0 = total;
                recs += set.files[f].n;
                total += x.fast ? x.map.n_bytes : x.slow.seq.size();
            }
            set.n_bases = total;
            // Page-locking gigabytes costs more than it saves when a set crosses PCIe once per GPU (measured: 1.6 s
            // to pin 3 GB against 0.1 s to fill it): pageable memory on transparent huge pages by default, the
            // driver stages the H2D copies.  COMMET_NXN_PINNED=1 pins (cudaHostAlloc) instead.
            static const bool want_pinned = getenv("COMMET_NXN_PINNED") && atoi(getenv("COMMET_NXN_PINNED")) != 0;
            if (want_pinned) set.bases = static_cast<uint8_t *>(commet_host_alloc(total + 64));
            set.pinned = set.bases != nullptr;
            if (!set.bases) {
                void *p = nullptr;
                const size_t bytes = (total + 64 + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
                if (posix_memalign(&p, 2u << 20, bytes) != 0) die("out of host memory");
                madvise(p, bytes, MADV_HUGEPAGE);
                set.bases = static_cast<uint8_t *>(p);
            }
            set.offs.assign(recs + 1, 0);
            set.offs[recs] = total;
        }
        phase_done("alloc");
        parallel_for(items.size(), [&](size_t it) {
            Src &x = src[items[it].src];
            ReadSetData &set = d.sets[x.s];
            const uint64_t first = set.files[x.f].first;
            if (x.fast) {
                uint64_t pos = x.base0, rec = first;
                for (size_t c = 0; c < items[it].chunk; c++) { pos += x.map.bytes[c]; rec += x.map.records[c]; }
                x.map.pass<true>(items[it].chunk, set.bases, pos, set.offs.data(), rec);
            } else {
                if (!x.slow.seq.empty()) memcpy(set.bases + x.base0, x.slow.seq.data(), x.slow.seq.size());
                for (uint64_t r = 0; r < x.slow.nb_reads; r++) set.offs[first + r] = x.base0 + x.slow.off[r];
            }
        });
        for (Src &x : src) x.map.close();
    }
    phase_done("parse");
    if (parse_only) {
        std::cerr << "parse: " << phases.back().second << " s\n";
        ctx_thread.join();
        return 0;
    }
    // ---- devices: the contexts were being created while the files were loaded --------------------------------
    ctx_thread.join();
    const int use = (int)d.workers.size();
    for (Driver::Worker &w : d.workers) w.staged.assign(d.sets.size(), nullptr);
    d.say("commet_nxn: " + std::to_string(d.sets.size()) + " sets, k=" + std::to_string(opt.k) + " t=" + std::to_string(opt.t) +
          ", " + std::to_string(use) + " GPU(s)");
    phase_done("contexts");
    // ---- filtering, Commet.py:103-121 + src/filter_reads.cpp -----------------------------------------------------
    {
        if (with_bv) {
            for (ReadSetData &set : d.sets)
                for (SetFileInfo &f : set.files) {
                    BitVec b;
                    b.read(f.bv_path);
                    d.store.put(f.bv_path, b);
                }
        } else {
            std::vector<std::thread> th;
            for (int g = 0; g < use; g++)
                th.emplace_back([&, g]() {
                    Driver::Worker &w = d.workers[g];
                    for (size_t s = g; s < d.sets.size(); s += use) {
                        ReadSetData &set = d.sets[s];
                        commet_reads *rs = d.stage(w, (int)s);
                        // local_m = m / len(files) is passed as a Python float ("4500.0"); filter_reads reads it with atoi
                        long cap_opt = opt.m >= 0 ? (long)atoi(py_float((double)opt.m / (double)set.files.size()).c_str()) : -1;
                        for (SetFileInfo &f : set.files) {
                            // the reference loop ends at the first empty read (filter_reads.cpp:188)
                            uint64_t n_eff = f.n;
                            for (uint64_t r = 0; r < f.n; r++)
                                if (set.offs[f.first + r + 1] == set.offs[f.first + r]) { n_eff = r; break; }
                            long max_reads = cap_opt == -1 ? (long)f.n : cap_opt;
                            std::vector<uint8_t> part(n_eff / 8 + 1, 0);
                            uint64_t counters[4] = {0, 0, 0, 0};
                            if (commet_filter_reads_range(w.ctx, rs, f.first, n_eff, opt.l, opt.n >= 0 ? opt.n : -1,
                                                          (float)atof(opt.e_text.c_str()), max_reads, part.data(), counters) != 0)
                                die(commet_last_error());
                            BitVec bv;
                            bv.init_true(f.n);
                            for (uint64_t r = 0; r < n_eff; r++)
                                if (!((part[r / 8] >> (r % 8)) & 1u)) bv.unset(r);
                            if ((long)counters[3] >= max_reads)
                                for (uint64_t r = n_eff; r < f.n; r++) bv.unset(r);
                            std::stringstream comment;             // src/filter_reads.cpp:160-176
                            comment << "----------------\nReference file\n";
                            size_t pos = f.path.rfind("/");
                            if (pos > 0 && pos < f.path.size()) comment << "  " << f.path.substr(pos + 1) << "\n";
                            else comment << "  " << f.path << "\n";
                            comment << "Filter Options\n";
                            comment << "  min read size     : " << opt.l << "\n";
                            if (opt.n < 0) comment << "  max number of N   : infinite\n";
                            else comment << "  max number of N   : " << opt.n << "\n";
                            comment << "  min shannon index : " << (float)atof(opt.e_text.c_str()) << "\n";
                            bv.comment = comment.str();
                            bv.write(f.bv_path);
                            d.store.put(f.bv_path, bv);
                            d.say("  filter " + f.path + ": " + std::to_string(counters[3]) + " / " + std::to_string(f.n) + " reads selected");
                        }
                    }
                });
            for (auto &t : th) t.join();
        }
    }

    phase_done("stage_and_filter");
    // ---- the N^2-1 rounds, Commet.py:186-240 -----------------------------------------------------------------------
    const int N = (int)d.sets.size();
    auto filter_bvs = [&](int s) {
        std::vector<std::string> v;
        for (const SetFileInfo &f : d.sets[s].files) v.push_back(f.bv_path);
        return v;
    };
    auto in_bvs = [&](int s, int other) {          // generate_A_File_Of_File_Index_WRT_A_Set
        std::vector<std::string> v;
        for (const SetFileInfo &f : d.sets[s].files) v.push_back(opt.out_dir + basename_of(f.path) + "_in_" + basename_of(d.sets[other].name) + ".bv");
        return v;
    };
    std::vector<std::vector<uint64_t>> shared(N, std::vector<uint64_t>(N, 0));
    std::mutex res_mu;

    // kind 0: "all in Si" for the query sets [j, j + nj) ; kind 1: the two refinement rounds of pair (i, j).
    // With several GPUs an "all in Si" round is cut into slices of a few query sets (Si is indexed once per slice):
    // one call over all N-1-i query sets would be the critical path of the whole run.
    struct Task { int kind, i, j, nj; };
    std::deque<Task> ready;
    std::mutex q_mu;
    std::condition_variable q_cv;
    int outstanding = 0;
    // vectors shared by name between rounds are only safe to compute out of order when no two files (or sets) share
    // a name; otherwise every round runs in the reference's order on one GPU
    bool unique_names = true;
    {
        std::set<std::string> seen;
        for (const ReadSetData &s : d.sets) {
            if (!seen.insert("set:" + s.name).second) unique_names = false;
            for (const SetFileInfo &f : s.files)
                if (!seen.insert("file:" + basename_of(f.path)).second) unique_names = false;
        }
    }
    auto run_all_in = [&](Driver::Worker &w, int i, int j0, int nj) {
        Round r;
        r.index_set = i;
        r.index_bvs = filter_bvs(i);
        for (int j = j0; j < j0 + nj; j++) { r.query_sets.push_back(j); r.query_bvs.push_back(filter_bvs(j)); }
        r.write_logs = false;                     // overwritten by the third round of every pair
        RoundOut o = d.run_round(w, r);
        d.say("  sets " + std::to_string(j0) + ".." + std::to_string(j0 + nj - 1) + " in " + d.sets[i].name + ": " + std::to_string(o.chunks) +
              " chunk(s), " + std::to_string(o.total_s * 1e3) + " ms on GPU " + std::to_string(w.device));
    };
    auto run_pair = [&](Driver::Worker &w, int i, int j) {
        Round b;                                   // Si in (Sj in Si)
        b.index_set = j;
        b.index_bvs = in_bvs(j, i);
        b.query_sets = {i};
        b.query_bvs = {filter_bvs(i)};
        RoundOut ob = d.run_round(w, b);
        Round c;                                   // Sj in (Si in (Sj in Si)): overwrites <Sj files>_in_<Si>.bv
        c.index_set = i;
        c.index_bvs = in_bvs(i, j);
        c.query_sets = {j};
        c.query_bvs = {filter_bvs(j)};
        RoundOut oc = d.run_round(w, c);
        {
            std::lock_guard<std::mutex> g(res_mu);
            shared[i][j] = ob.ones[0];
            shared[j][i] = oc.ones[0];
        }
        d.say("  " + d.sets[i].name + " x " + d.sets[j].name + ": " + std::to_string(ob.ones[0]) + " / " + std::to_string(oc.ones[0]) +
              " shared reads, " + std::to_string((ob.total_s + oc.total_s) * 1e3) + " ms on GPU " + std::to_string(w.device));
    };

    if (!unique_names || d.workers.size() == 1) {
        Driver::Worker &w = d.workers[0];
        for (int i = 0; i + 1 < N; i++) {
            run_all_in(w, i, i + 1, N - 1 - i);
            for (int j = i + 1; j < N; j++) run_pair(w, i, j);
        }
    } else {
        const int slice = std::max(1, (N + 2) / 4);
        for (int i = 0; i + 1 < N; i++)
            for (int j = i + 1; j < N; j += slice) ready.push_back({0, i, j, std::min(slice, N - j)});
        outstanding = (int)ready.size();
        std::vector<std::thread> th;
        for (size_t g = 0; g < d.workers.size(); g++)
            th.emplace_back([&, g]() {
                Driver::Worker &w = d.workers[g];
                while (true) {
                    Task t;
                    {
                        std::unique_lock<std::mutex> lk(q_mu);
                        q_cv.wait(lk, [&]() { return !ready.empty() || outstanding == 0; });
                        if (ready.empty()) return;
                        // refinement rounds first: they free no new work and keep the sets of a pair hot; take the
                        // largest "all in" (smallest i) otherwise
                        size_t pick = 0;
                        for (size_t x = 0; x < ready.size(); x++)
                            if (ready[x].kind == 1) { pick = x; break; }
                        t = ready[pick];
                        ready.erase(ready.begin() + (std::ptrdiff_t)pick);
                    }
                    if (t.kind == 0) run_all_in(w, t.i, t.j, t.nj);
                    else run_pair(w, t.i, t.j);
                    {
                        std::lock_guard<std::mutex> lk(q_mu);
                        if (t.kind == 0)
                            for (int j = t.j; j < t.j + t.nj; j++) { ready.push_back({1, t.i, j, 1}); outstanding++; }
                        outstanding--;
                    }
                    q_cv.notify_all();
                }
            });
        for (auto &t : th) t.join();
    }

    phase_done("rounds");
    // ---- write the vectors the reference leaves on disk ----------------------------------------------------------------
    for (auto &kv : d.store.v) {
        bool is_filter = false;
        for (const ReadSetData &s : d.sets)
            for (const SetFileInfo &f : s.files) is_filter |= f.bv_path == kv.first;
        if (!is_filter) kv.second.write(kv.first);
    }

    // ---- matrices, Commet.py:245-317 --------------------------------------------------------------------------------
    std::vector<uint64_t> totals(N, 0);
    for (int s = 0; s < N; s++)
        for (const SetFileInfo &f : d.sets[s].files) {
            BitVec b = d.store.get(f.bv_path);
            uint64_t ones = 0;                 // nb_one: popcount of ALL payload bytes, clamped (boolean_vector.h:244-270)
            if (commet_bv_popcount(d.workers[0].ctx, b.bytes.data(), b.n, &ones) != 0) die(commet_last_error());
            totals[s] += ones;
        }
    for (int s = 0; s < N; s++) shared[s][s] = totals[s];
    std::string head;
    for (const ReadSetData &s : d.sets) head += ";" + s.name;
    head += "\n";
    std::ofstream plain((opt.out_dir + "matrix_plain.csv").c_str()), pct((opt.out_dir + "matrix_percentage.csv").c_str()),
        norm((opt.out_dir + "matrix_normalized.csv").c_str());
    plain << head;
    pct << head;
    norm << head;
    for (int i = 0; i < N; i++) {
        plain << d.sets[i].name;
        pct << d.sets[i].name;
        norm << d.sets[i].name;
        for (int j = 0; j < N; j++) {
            if (totals[i] == 0 || totals[i] + totals[j] == 0) die("ZeroDivisionError: set " + d.sets[i].name + " has no selected read");
            plain << ";" << shared[i][j];
            pct << ";" << py_float((double)(100 * shared[i][j]) / (double)totals[i]);
            norm << ";" << py_float((double)(100 * (shared[i][j] + shared[j][i])) / (double)(totals[i] + totals[j]));
        }
        plain << "\n";
        pct << "\n";
        norm << "\n";
    }
    plain.close();
    pct.close();
    norm.close();

    phase_done("vectors_and_matrices");
    if (!opt.report.empty()) {
        std::ofstream rep(opt.report.c_str());
        uint64_t reads = 0, bases = 0;
        for (const ReadSetData &s : d.sets) { reads += s.n_records(); bases += s.n_bases; }
        rep << "{\"sets\": " << N << ", \"rounds\": " << (N * N - 1) << ", \"gpus\": " << d.workers.size() << ", \"reads\": " << reads
            << ", \"bases\": " << bases << ", \"k\": " << opt.k << ", \"t\": " << opt.t << ", \"seconds_at_end_of\": {";
        for (size_t i = 0; i < phases.size(); i++) rep << (i ? ", " : "") << "\"" << phases[i].first << "\": " << phases[i].second;
        rep << "}}\n";
    }
    d.say("commet_nxn: done in " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count()) + " s; matrices in " + opt.out_dir);
    // every output is on disk: leave without tearing down gigabytes of device and host allocations one by one
    std::cout.flush();
    std::cerr.flush();
    if (const char *e = getenv("COMMET_NXN_EXIT")) {          // A/B of the ways out (diagnostic)
        if (!strcmp(e, "return")) return 0;
        if (!strcmp(e, "destroy")) {
            for (Driver::Worker &w : d.workers) commet_ctx_destroy(w.ctx);
            _exit(0);
        }
    }
    _exit(0);
}
