// compare_reads -- drop-in for the reference tool of the same name (src/compare_reads.cpp): the three passes of a
// full A/B comparison behind their own argv.  B in A; then A in (B in A), written as <A file>_in_<B>.bv; then
// B in (A in (B in A)), written as <B file>_in_<A>.bv.  Same kernels and the same C-ABI call as
// `index_and_search -f`; what differs is the text on stdout, no .log files, and which intermediate vector is saved.
// There is no CPU path.
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <map>
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
    std::cerr << "\ncompare_reads, version " << version << "\n";
    std::cerr << "Usage : ./compare_reads -i <file> -s <file> [options]\n";
    std::cerr << "Mandatory:\n";
    std::cerr << "\t -i <file>: A file containing the list of files to index (comma separated) - MANDATORY\n";
    std::cerr << "\t            Each line of the file corresponds to a set of files (comma separated)\n";
    std::cerr << "\t -s <file>: A file containing the list of file sets to search - MANDATORY\n";
    std::cerr << "\t            Each line of the file corresponds to a set of files (comma separated)\n";
    std::cerr << "Options:\n";
    std::cerr << "\t -l </.../>: ABSOLUTE path to log folder\n";
    std::cerr << "\t -o </.../>: ABSOLUTE path to output folder\n";
    std::cerr << "\t -k <value>: Size of k-mers (value of k). [default=32]\n";
    std::cerr << "\t -t <value>: Number of shared k-mers. [default=2]\n";
    std::cerr << "\t -h: Prints this message and exit\n";
    std::cerr << "\t -v: Prints the version number and exit\n";
}

static void print_times(const PassResult &r)
{
    std::cout << "Index  time: " << (float)r.index_s << " s\n";
    std::cout << "Search time: " << (float)r.search_s << " s\n";
    std::cout << "Total  time: " << (float)r.total_s << " s\n";
    std::cout << "[indexed " << r.indexed << ", searched " << r.searched[0] << ", shared " << r.shared[0] << "]";
}

// The reference repeats index+search `while (nb_indexed_reads < nb_reads_to_index)` (src/compare_reads.cpp:248):
// a read fetched and lost at a chunk boundary (index_reads.h:48-49,60) is never counted, so with more than one
// chunk that loop never ends.  Here the pass ends when the set is exhausted, as index_and_search's own loop does
// (src/index_and_search.cpp:255), and the divergence is reported.
static void warn_if_reference_would_hang(const PassResult &r, const ReadSet &index)
{
    if (r.indexed < index.n_valid())
        std::cerr << "compare_reads: " << index.n_valid() - r.indexed << " read(s) of {" << index.nickname
                  << "} were fetched and lost at chunk boundaries; the reference tool does not terminate on this input\n";
}

int main(int argc, char **argv)
{
    std::string A_file_list, B_file_list;
    int kmer_size = 33;
    int min_hits = 2;
    uint64_t max_kmer = commet_max_kmer(kmer_size);
    std::string log_path = ".";
    std::string out_path = ".";

    // ---- argv, src/compare_reads.cpp:82-164 -----------------------------------
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
            if (!A_file_list.empty()) std::cerr << "A files already given (-i) -> ignore";
            else A_file_list = argv[arg_pos];
        } else if (flag == "-s") {
            need_arg();
            if (!B_file_list.empty()) std::cerr << "B files already given (-s) -> ignore";
            else B_file_list = argv[arg_pos];
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
        } else if (flag == "-h") {
            print_usage();
            return 0;
        } else if (flag == "-v") {
            std::cout << "\ncompare_reads version " << version << "\n";
            return 0;
        } else {
            std::cerr << "Unknown option " << flag << "\n";
            print_usage();
            return 0;
        }
        arg_pos++;
    }

    ensure_dir(log_path);          // :169-183 (log first, then out)
    ensure_dir(out_path);

    // ---- sets, :188-224: more than one set is a warning, the first one is kept ----
    std::map<std::string, SetSpec> A_specs = read_sets(A_file_list);
    if (A_specs.size() != 1) std::cerr << "Only one set of files is allowed for A -> keep first set only\n";
    if (A_specs.empty()) return 1;      // the reference dereferences begin() of an empty map here
    ReadSet A;
    A.nickname = A_specs.begin()->first;
    load_set(A, A_specs.begin()->second);
    std::map<std::string, SetSpec> B_specs = read_sets(B_file_list);
    if (B_specs.size() != 1) std::cerr << "Only one set of files is allowed for B -> keep first set only\n";
    if (B_specs.empty()) return 1;
    ReadSet B;
    B.nickname = B_specs.begin()->first;
    load_set(B, B_specs.begin()->second);
    A.build_stream(true);
    B.build_stream(true);
    const uint64_t nb_reads_A = A.n_valid(), nb_reads_B = B.n_valid();

    Engine ctx;
    if (!ctx.open(A.bases.size() + B.bases.size())) {
        std::cerr << "compare_reads: " << commet_last_error() << "\n";
        return 1;
    }
    auto banner = [](const std::string &text) {
        std::cout << "\n------------------------------------------------------------------\n";
        std::cout << text << "\n";
        std::cout << "------------------------------------------------------------------\n";
    };
    const std::string &a = A.nickname, &b = B.nickname;

    // ---- B in A, :238-264 ------------------------------------------------------
    banner("finding reads from {" + b + "} present in raw {" + a + "}");
    std::vector<ReadSet *> qb{&B};
    PassResult r1 = run_pass(ctx, kmer_size, min_hits, max_kmer, A, qb, false, "compare_reads");
    warn_if_reference_would_hang(r1, A);
    B.apply_out_as_input();
    B.build_stream(true);
    print_times(r1);
    std::cout << "\n";

    // ---- A in (B in A), :269-298 -------------------------------------------------
    banner("finding reads from {" + a + "} present in raw {" + b + "} present in raw {" + a + "}");
    std::vector<ReadSet *> qa{&A};
    PassResult r2 = run_pass(ctx, kmer_size, min_hits, max_kmer, B, qa, false, "compare_reads");
    warn_if_reference_would_hang(r2, B);
    A.save_bv(out_path, b);
    A.apply_out_as_input();
    A.build_stream(true);
    print_times(r2);
    std::cout << " " << 100 * (float)r2.shared[0] / (float)nb_reads_A << "%\n";

    // ---- B in (A in (B in A)), :303-333 -------------------------------------------
    banner("finding reads from {" + b + "} present in raw {" + a + "} present in raw {" + b + "} present in raw {" + a + "}");
    PassResult r3 = run_pass(ctx, kmer_size, min_hits, max_kmer, A, qb, false, "compare_reads");
    warn_if_reference_would_hang(r3, A);
    B.save_bv(out_path, a);
    print_times(r3);
    std::cout << " " << 100 * (float)r3.shared[0] / (float)nb_reads_B << "%\n";

    ctx.close();
    return 0;
}
