// index_and_search -- drop-in for the reference tool of the same name
// (src/index_and_search.cpp): same argv, same fof format, same stdout text,
// same <file>_in_<set>.bv and <query>_in_<index>.log outputs.  The host side
// parses the read files and drives the C-ABI (include/commet_b200.h); the
// chunk loop, indexing and search run on the GPU.  There is no CPU path.
#include <sys/stat.h>
#include <sys/types.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "commet_b200.h"
#include "pass.hpp"
#include "read_set.hpp"
#include "set_parser.hpp"

using namespace commet_host;

static const std::string version = "2.1";

static void print_usage()
{
    std::cerr << "\nindex_and_search, version " << version << "\n";
    std::cerr << "Usage : ./index_and_search -i <file> -s <file> [options]\n";
    std::cerr << "Mandatory:\n";
    std::cerr << "\t -i <file>: A file containing the list of files to index - MANDATORY\n";
    std::cerr << "\t -s <file>: A file containing the list of files to search - MANDATORY\n";
    std::cerr << "\t            Each line of the file corresponds to a set of files to search\n";
    std::cerr << "Options:\n";
    std::cerr << "\t -l </.../>: ABSOLUTE path to log folder\n";
    std::cerr << "\t -o </.../>: ABSOLUTE path to output folder\n";
    std::cerr << "\t -k <value>: Size of k-mers (value of k). [default=33]\n";
    std::cerr << "\t -t <value>: Number of shared k-mers. [default=2]\n";
    std::cerr << "\t -f: Full comparison of index set and the first searched set [default=false]\n";
    std::cerr << "\t -h: Prints this message\n";
    std::cerr << "\t -v: Prints the version number\n";
}

template <class Stream>
static void print_times(Stream &o, const PassResult &r, size_t s)
{
    o << "Index  time: " << (float)r.index_s << " s\n";
    o << "Search time: " << (float)r.search_s << " s\n";
    o << "Total  time: " << (float)r.total_s << " s\n";
    o << "[indexed " << r.indexed << ", searched " << r.searched[s] << ", shared " << r.shared[s] << "]\n";
}

int main(int argc, char **argv)
{
    std::string search_file_list, index_file_list;
    int kmer_size = 33;
    int min_hits = 2;
    uint64_t max_kmer = commet_max_kmer(kmer_size);
    std::string log_path = ".";
    std::string out_path = ".";
    bool full = false;

    // ---- argv, src/index_and_search.cpp:91-172 -------------------------------
    if (argc == 1) {
        print_usage();
        return 0;
    }
    int arg_pos = 1;
    auto need_arg = [&]() {
        arg_pos++;
        if (arg_pos >= argc) {
            std::cerr << "Error, flag " << argv[arg_pos - 1] << " needs an argument\n";
            print_usage();
            exit(1);
        }
    };
    while (arg_pos < argc) {
        std::string flag = argv[arg_pos];
        if (flag == "-i") {
            need_arg();
            if (!index_file_list.empty()) std::cerr << "index files already given (-i) -> ignore";
            else index_file_list = argv[arg_pos];
        } else if (flag == "-s") {
            need_arg();
            if (!search_file_list.empty()) std::cerr << "search files already given (-s) -> ignore";
            else search_file_list = argv[arg_pos];
        } else if (flag == "-l") {
            need_arg();
            log_path = argv[arg_pos];
        } else if (flag == "-o") {
            need_arg();
            out_path = argv[arg_pos];
        } else if (flag == "-k") {
            need_arg();
            kmer_size = atoi(argv[arg_pos]);
            max_kmer = (unsigned long)(1000000000.0 / pow(2, 33 - kmer_size));
            std::cout << "k-mer size (-k) = " << kmer_size << "\n";
        } else if (flag == "-t") {
            need_arg();
            min_hits = atoi(argv[arg_pos]);
            std::cout << "min hits (-t) = " << min_hits << "\n";
        } else if (flag == "-f") {
            full = true;
        } else if (flag == "-h") {
            print_usage();
            return 0;
        } else if (flag == "-v") {
            std::cout << "\nindex_and_search version " << version << "\n";
            return 0;
        } else {
            std::cerr << "Unknown option " << flag << "\n";
            print_usage();
            return 0;
        }
        arg_pos++;
    }

    ensure_dir(log_path);
    ensure_dir(out_path);
    Engine ctx;
    ctx.prewarm();                   // CUDA comes up while the read files are parsed
    ParseAhead ahead;
    ahead.start(search_file_list);   // ... and the search sets' FASTA files while the index set is loaded

    // ---- sets, src/index_and_search.cpp:196-234 ------------------------------
    std::map<std::string, SetSpec> index_specs = read_sets(index_file_list);
    if (index_specs.size() != 1) {
        std::cerr << "Only one set of files is allowed for indexing\n";
        exit(1);
    }
    ReadSet index_set;
    index_set.nickname = index_specs.begin()->first;
    load_set(index_set, index_specs.begin()->second);

    std::map<std::string, SetSpec> search_specs = read_sets(search_file_list);
    std::vector<std::unique_ptr<ReadSet>> search_sets;
    for (auto &kv : search_specs) {
        std::unique_ptr<ReadSet> s(new ReadSet);
        s->nickname = kv.first;
        load_set(*s, kv.second, &ahead);
        search_sets.push_back(std::move(s));
        if (full) break;
    }
    if (search_sets.empty()) {
        std::cerr << "No set of files to search\n";
        exit(1);
    }
    index_set.build_stream(full);
    for (auto &s : search_sets) s->build_stream(full);
    const uint64_t nb_reads_A = index_set.n_valid();
    const uint64_t nb_reads_B = search_sets[0]->n_valid();

    uint64_t total_bases = index_set.bases.size();
    for (auto &s : search_sets) total_bases += s->bases.size();
    ctx.mark("read sets parsed, valid-read streams built");
    if (!ctx.open(total_bases)) {
        std::cerr << "index_and_search: " << commet_last_error() << "\n";
        return 1;
    }

    // ---- first pass, :241-301 -------------------------------------------------
    std::vector<ReadSet *> queries;
    for (auto &s : search_sets) queries.push_back(s.get());
    PassResult r1 = run_pass(ctx, kmer_size, min_hits, max_kmer, index_set, queries, true);
    ctx.mark("first pass done");
    for (size_t s = 0; s < queries.size(); s++) {
        std::cout << "\n------------------------------------------------------------------\n";
        std::cout << "Reads from {" << queries[s]->nickname << "} present in raw {" << index_set.nickname << "}\n";
        std::cout << "------------------------------------------------------------------\n";
        print_times(std::cout, r1, s);
        std::string fname = log_path + "/" + queries[s]->nickname + "_in_" + index_set.nickname + ".log";
        std::ofstream log_file(fname.c_str());
        if (!log_file.good()) {
            std::cerr << "Cannot open log file : " << fname << "\n";
            exit(1);
        }
        print_times(log_file, r1, s);
    }

    // ---- -f: second and third pass, :304-391 ---------------------------------
    if (full) {
        ReadSet &B = *search_sets[0];
        std::string log_name = log_path + "/" + index_set.nickname + "_in_" + B.nickname + ".log";
        std::ofstream log_file(log_name.c_str());
        if (!log_file.good()) {
            std::cerr << "Cannot open log file " << log_name << " -> exit\n";
            exit(1);
        }
        B.apply_out_as_input();
        B.build_stream(true);
        std::cout << "\n------------------------------------------------------------------\n";
        std::cout << "finding reads from {" << index_set.nickname << "} present in {raw {" << B.nickname
                  << "} present in raw {" << index_set.nickname << "}}\n";
        std::cout << "------------------------------------------------------------------\n";
        std::vector<ReadSet *> qa{&index_set};
        PassResult r2 = run_pass(ctx, kmer_size, min_hits, max_kmer, B, qa, false);
        index_set.save_bv(out_path, B.nickname);
        index_set.apply_out_as_input();
        index_set.build_stream(true);
        print_times(std::cout, r2, 0);
        std::cout << 100 * (float)r2.shared[0] / (float)nb_reads_A << "%\n";
        print_times(log_file, r2, 0);
        log_file << 100 * (float)r2.shared[0] / (float)nb_reads_A << "%\n";
        log_file.close();

        log_name = log_path + "/" + B.nickname + "_in_" + index_set.nickname + ".log";
        log_file.open(log_name.c_str());
        if (!log_file.good()) {
            std::cerr << "Cannot open log file " << log_name << " -> exit\n";
            exit(1);
        }
        std::cout << "\n------------------------------------------------------------------\n";
        std::cout << "finding reads from {" << B.nickname << "} present in {raw {" << index_set.nickname
                  << "} present in {raw {" << B.nickname << "} present in raw {" << index_set.nickname << "}}}\n";
        std::cout << "------------------------------------------------------------------\n";
        std::vector<ReadSet *> qb{&B};
        PassResult r3 = run_pass(ctx, kmer_size, min_hits, max_kmer, index_set, qb, false);
        B.save_bv(out_path, index_set.nickname);
        print_times(std::cout, r3, 0);
        std::cout << 100 * (float)r3.shared[0] / (float)nb_reads_B << "%\n";
        print_times(log_file, r3, 0);
        log_file << 100 * (float)r3.shared[0] / (float)nb_reads_B << "%\n";
        log_file.close();
    }

    ctx.mark("passes done");
    // ---- outputs, :397-399 ------------------------------------------------------
    for (auto &s : search_sets) s->save_bv(out_path, index_specs.begin()->first);
    ctx.mark("vectors written");
    // every output is on disk: the process ends without tearing the CUDA context down (the driver reclaims the device
    // memory of a dead process; an orderly teardown of a 15 GB context was measured at 30-440 ms)
    std::cout.flush();
    std::cerr.flush();
    std::_Exit(0);
}
