// filter_reads -- drop-in for src/filter_reads.cpp: same argv, same comment
// block, same stdout statistics, same .bv output.  The per-read selection
// (length, N count, Shannon index, -m cap) runs on the GPU through
// commet_filter_reads; there is no CPU path.
#include <limits.h>

#include <chrono>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

#include "bv.hpp"
#include "commet_b200.h"
#include "readers.hpp"

using namespace commet_host;

static const std::string version = "2.1";

static void print_usage()
{
    std::cout << "\nfilter_reads v" << version << "\n";
    std::cout << "Usage:\n\t./filter_reads <input_file> [options]\n";
    std::cout << "Mandatory:\n";
    std::cout << "\t<input_file>\t: file containing reads, in fasta or fastq format, gzipped or not\n";
    std::cout << "Options:\n";
    std::cout << "\t -o string\t: file where the boolean vector will be written [default=input_file.bv]\n";
    std::cout << "\t -l int\t\t: minimal length a read should have to be kept. [default=0]\n";
    std::cout << "\t -n int\t\t: maximal number of Ns a read should contain to be kept. [default=any]\n";
    std::cout << "\t -e float\t: minimal Shannon index a read should have to be kept. [default=0]\n";
    std::cout << "\t -m int\t\t: maximum number of selected reads [default=all]\n";
    std::cout << "\t -c string\t: the given string will be written in the header of the output file. [default=command line]\n";
    std::cout << "\t -h\t\t: prints this help\n";
    std::cout << "\t -v\t\t: prints the version number.\n\n";
}

int main(int argc, char **argv)
{
    auto begin = std::chrono::steady_clock::now();
    std::string input_file_name, output_file_name;
    int min_size = 0;
    int max_N = INT_MAX;
    float min_shannon = 0.0;
    std::stringstream comment;
    long max_reads = -1;

    // ---- argv, src/filter_reads.cpp:63-104 (a flag's missing value is read as the reference reads it: argv[argc]) --
    int arg_pos = 1;
    auto value = [&]() -> const char * {
        arg_pos++;
        if (arg_pos >= argc) {
            std::cerr << "Error, flag " << argv[arg_pos - 1] << " needs an argument\n";
            exit(1);
        }
        return argv[arg_pos];
    };
    while (arg_pos < argc) {
        std::string flag = argv[arg_pos];
        if (flag.empty() || flag[0] != '-') {
            if (input_file_name.empty()) input_file_name = flag;
            else if (output_file_name.empty()) output_file_name = flag;
            else std::cout << "The mandatory files are already set, unknown file " << flag << " -> ignore\n";
        } else if (flag == "-o") {
            output_file_name = value();
        } else if (flag == "-l") {
            min_size = atoi(value());
        } else if (flag == "-n") {
            max_N = atoi(value());
        } else if (flag == "-m") {
            max_reads = atoi(value());
        } else if (flag == "-e") {
            min_shannon = atof(value());
        } else if (flag == "-c") {
            comment << value() << "\n";
        } else if (flag == "-h") {
            print_usage();
            return 0;
        } else if (flag == "-v") {
            std::cout << "\nfilter_reads version " << version << "\n";
            return 0;
        } else {
            std::cerr << "Unknown option " << flag << "\n";
            print_usage();
            return 1;
        }
        arg_pos++;
    }
    if (input_file_name.empty()) {
        std::cerr << "Error: An input file name is needed -> exit\n";
        print_usage();
        return 0;
    }
    std::string output_message;
    if (output_file_name.empty()) {
        output_message = "No output file name given, results will be written in " + input_file_name + ".bv\n";
        output_file_name = input_file_name + ".bv";
    }

    // ---- read the file (format sniffing of :121-154) ---------------------------
    ParsedFile pf;
    {
        std::ifstream probe(input_file_name.c_str());
        if (!probe.good()) {
            std::cerr << "Cannot open file " << input_file_name << " -> quit\n";
            return 1;
        }
    }
    if (!parse_reads_file(input_file_name, pf, " -> quit\n")) exit(1);

    // ---- comment block, :160-176 ------------------------------------------------
    comment << "----------------\n";
    comment << "Reference file\n";
    size_t pos = input_file_name.rfind("/");
    if (pos > 0 && pos < input_file_name.size()) comment << "  " << input_file_name.substr(pos + 1) << "\n";
    else comment << "  " << input_file_name << "\n";
    comment << "Filter Options\n";
    comment << "  min read size     : " << min_size << "\n";
    if (max_N == INT_MAX) comment << "  max number of N   : infinite\n";
    else comment << "  max number of N   : " << max_N << "\n";
    comment << "  min shannon index : " << min_shannon << "\n";

    // ---- selection on the GPU, :181-205 -------------------------------------------
    // The reference loop ends at the first empty read (:188): later reads are never examined.
    uint64_t n = pf.nb_reads, n_eff = n;
    for (uint64_t r = 0; r < n; r++)
        if (pf.off[r + 1] == pf.off[r]) { n_eff = r; break; }
    if (max_reads == -1) max_reads = (long)n;

    commet_ctx *ctx = nullptr;
    if (commet_ctx_create(0, &ctx) != 0) {
        std::cerr << "filter_reads: " << commet_last_error() << "\n";
        return 1;
    }
    BitVec bv;
    bv.init_true(n);
    uint64_t counters[4] = {0, 0, 0, 0};
    {
        std::vector<uint8_t> part(n_eff / 8 + 1, 0);
        static const uint8_t none = 0;
        // a cap that is never reached inside the examined prefix behaves like "no cap" there
        long cap = max_reads;
        if (commet_filter_reads(ctx, pf.seq.empty() ? &none : pf.seq.data(), pf.off.data(), n_eff, min_size,
                                max_N == INT_MAX ? -1 : (max_N < 0 ? -2 : max_N), min_shannon, cap, part.data(), counters) != 0) {
            std::cerr << "filter_reads: " << commet_last_error() << "\n";
            return 1;
        }
        for (uint64_t r = 0; r < n_eff; r++)
            if (!((part[r / 8] >> (r % 8)) & 1u)) bv.unset(r);
        // untag_last_reads (read_file.h:76-81) also clears the never-examined tail once the cap is reached
        if ((long)counters[3] >= max_reads)
            for (uint64_t r = n_eff; r < n; r++) bv.unset(r);
    }
    commet_ctx_destroy(ctx);
    bv.comment = comment.str();
    bv.write(output_file_name);

    std::cout << "Length filter [" << min_size << "]: " << counters[0] << " reads removed\n";
    if (max_N == INT_MAX) std::cout << "Number of N filter [infinite]: " << counters[1] << " reads removed\n";
    else std::cout << "Number of N filter [" << max_N << "]: " << counters[1] << " reads removed\n";
    std::cout << "Shannon filter [" << min_shannon << "]: " << counters[2] << " reads removed\n";
    std::cout << "Number of selected reads = " << counters[3] << "\n";
    if (!output_message.empty()) std::cout << output_message;
    std::cout << "Total  time : "
              << (float)std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count() << " s\n";
    return 0;
}
