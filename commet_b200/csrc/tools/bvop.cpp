// bvop -- drop-in for src/bvop.cpp: AND / OR / AND-NOT / NOT of .bv files,
// `-i` information line parsed by Commet.py:257,269, `-p` output.  The
// byte-wise operators and the popcount run on the GPU (commet_bvop,
// commet_bv_popcount); there is no CPU path.
#include <iostream>
#include <string>

#include "bv.hpp"
#include "commet_b200.h"

using namespace commet_host;

static const std::string version = "2.1";

static void print_usage()
{
    std::cout << "\nbvop, version " << version << "\n";
    std::cout << "Usage : ./bvop <file1.bv> [options]\n";
    std::cout << "Mandatory:\n";
    std::cout << "\t<file1.bv>\t: file containing a boolean vector\n";
    std::cout << "Options:\n";
    std::cout << "\t -n             : performs NOT on file1.bv\n";
    std::cout << "\t -a <file2.bv>  : performs file1.bv AND file2.bv\n";
    std::cout << "\t -o <file2.bv>  : performs file1.bv OR file2.bv\n";
    std::cout << "\t -d <file2.bv>  : performs file1.bv AND (NOT file2.bv)\n";
    std::cout << "\t -p <output.bv> : print result in file output.bv [Default=stdout]\n";
    std::cout << "\t -i             : print information about file1.bv\n";
    std::cout << "\t -h             : Prints this message and exit\n";
    std::cout << "\t -v             : Prints the version number and exit\n";
}

static void die_gpu()
{
    std::cerr << "bvop: " << commet_last_error() << "\n";
    exit(1);
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::cerr << "A boolean vector file must be provided, see usage\n";
        print_usage();
        return 1;
    }
    std::string file_name1, file_name2, output_file_name;
    bool print = false, print_info = false;
    char op = 'u';
    int i = 1;
    auto value = [&]() -> const char * {
        i++;
        if (i >= argc) {
            std::cerr << "Error, flag " << argv[i - 1] << " needs an argument\n";
            exit(1);
        }
        return argv[i];
    };
    while (i < argc) {                                   // src/bvop.cpp:73-124
        if (argv[i][0] == '-') {
            switch (argv[i][1]) {
            case 'a': file_name2 = value(); op = 'a'; break;
            case 'o': file_name2 = value(); op = 'o'; break;
            case 'd': file_name2 = value(); op = 'd'; break;
            case 'n': op = 'n'; break;
            case 'p': output_file_name = value(); print = true; break;
            case 'i': print_info = true; break;
            case 'v': std::cout << "compare_reads version " << version << "\n"; return 0;
            case 'h':
            default: print_usage(); return 0;
            }
        } else {
            if (file_name1.empty()) {
                file_name1 = argv[i];
            } else {
                std::cerr << "One input file is mandatory\n";
                print_usage();
                return 0;
            }
        }
        i++;
    }

    BitVec bv1;
    bv1.read(file_name1);
    // everything the host can check comes before the device is touched: the second vector is read and sized first
    // (boolean_vector.h:420-423), and an invocation with nothing to compute never creates a context
    BitVec bv2;
    const bool binary = op == 'a' || op == 'o' || op == 'd';
    if (binary) {
        bv2.read(file_name2);
        if (bv2.n != bv1.n) {
            std::cerr << "Error: the two vectors are not the same size -> exit\n";
            exit(1);
        }
    }
    commet_ctx *ctx = nullptr;
    if ((binary || op == 'n' || print_info) && commet_ctx_create(0, &ctx) != 0) die_gpu();

    std::string comment;
    bool do_nothing = false;
    if (binary) {
        int code = op == 'a' ? COMMET_BV_AND : op == 'o' ? COMMET_BV_OR : COMMET_BV_ANDNOT;
        if (commet_bvop(ctx, code, bv1.bytes.data(), bv2.bytes.data(), bv1.bytes.data(), bv1.bytes.size()) != 0) die_gpu();
        comment = file_name1 + (op == 'a' ? " AND " : op == 'o' ? " OR " : " AND (NOT ") + file_name2 +
                  (op == 'd' ? ")\n" : "\n");
    } else if (op == 'n') {
        if (commet_bvop(ctx, COMMET_BV_NOT, bv1.bytes.data(), nullptr, bv1.bytes.data(), bv1.bytes.size()) != 0) die_gpu();
        comment = "NOT " + file_name1 + "\n";
    } else {
        do_nothing = true;
    }

    if (print_info) {                                    // src/bvop.cpp:155-160
        uint64_t ones = 0;
        if (commet_bv_popcount(ctx, bv1.bytes.data(), bv1.n, &ones) != 0) die_gpu();
        std::cout << bv1.comment;
        std::cout << "\nReads:\n";
        std::cout << "  " << ones << " / " << bv1.n << " reads selected\n";
    }
    if (ctx) commet_ctx_destroy(ctx);
    if (do_nothing) return 0;

    bv1.comment = comment;
    if (print) bv1.write(output_file_name);
    else bv1.print_stdout();
    return 0;
}
