// extract_reads -- drop-in for src/extract_reads.cpp: writes the records of a read file whose bit is set in a
// boolean vector (the .bv files index_and_search / filter_reads / bvop produce), as FASTA/FASTQ text, gzip
// output when the input is gzipped.  Pure host I/O: the file is parsed ONCE from memory instead of
// the reference's counting pass + getline/gzgetc pass, and the record text is the reference's get_data():
//   FASTA plain : header line + every NON-EMPTY sequence line, each followed by '\n'   (fasta_file.h:156-176)
//   FASTA gzip  : header line, then the bytes up to the next '>' verbatim              (fasta_file.h:415-435)
//   FASTQ plain : the record's four non-empty lines, each followed by '\n'             (fastq_file.h:152-196)
//   FASTQ gzip  : four consecutive lines, no blank-line skipping                       (fastq_file.h:467-521)
// Extraction stops at the first selected record with an empty sequence (the reference's end-of-stream
// sentinel, extract_reads.cpp:153-156) and after popcount(bv) records (fasta_file.h:142).
#include <zlib.h>

#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "bv.hpp"
#include "readers.hpp"

using namespace commet_host;

static const std::string version = "2.1";

static void print_usage()
{
    std::cout << "\nextract_reads v" << version << "\n";
    std::cout << "Usage:\n\t./extract_reads <input_file> <bv_file> [options]\n";
    std::cout << "Mandatory:\n";
    std::cout << "\t<input_file>\t: file containing reads, in fasta or fastq format, gzipped or not\n";
    std::cout << "\t<bv_file>\t: associated boolean vector file\n";
    std::cout << "Options:\n";
    std::cout << "\t -o string: Output results in the given file [default=stdout]\n";
    std::cout << "\t -h: Prints this message and exit\n";
    std::cout << "\t -v: prints the version number.\n\n";
    exit(0);
}

namespace {

struct Line { const char *b, *e; bool nl; };     // [b, e) without the terminator; nl: a '\n' followed

struct Cursor {
    const char *p, *end;
    bool eof() const { return p >= end; }
    Line next()
    {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        Line l{p, nl ? nl : end, nl != nullptr};
        p = nl ? nl + 1 : end;
        return l;
    }
};

struct Sink {
    std::string buf;
    virtual void flush_buf() = 0;
    void put(const char *b, const char *e)
    {
        buf.append(b, e);
        if (buf.size() >= (8u << 20)) flush_buf();
    }
    void put(char c) { buf.push_back(c); }
    virtual ~Sink() {}
};
struct FileSink : Sink {
    FILE *f;
    explicit FileSink(FILE *f) : f(f) {}
    void flush_buf() override { if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
};
struct GzSink : Sink {
    gzFile g;
    explicit GzSink(gzFile g) : g(g) {}
    void flush_buf() override { if (!buf.empty() && g) gzwrite(g, buf.data(), (unsigned)buf.size()); buf.clear(); }
};

// Every extractor returns after the record budget (`left`) is used or an empty sequence is met.
void extract_fasta(const std::string &t, bool gz, const BitVec &bv, uint64_t nb_reads, uint64_t left, Sink &out)
{
    Cursor c{t.data(), t.data() + t.size()};
    uint64_t pos = 0;
    std::string rec;
    // records start at lines beginning with '>' (fasta_file.h:61-68 counts exactly those)
    while (!c.eof() && pos < nb_reads && left > 0) {
        Line h = c.next();
        if (h.e == h.b || *h.b != '>') {
            if (h.e == h.b) continue;                // blank line before the first record
            std::cerr << "Error in Fasta format !!\n";
            exit(1);
        }
        const bool take = bv.get(pos);
        size_t seq_len = 0;
        rec.clear();
        if (take) { rec.append(h.b, h.e); rec.push_back('\n'); }
        if (gz) {
            // verbatim up to the next '>' (any position: gzgetc loop of fasta_file.h:426-433)
            const char *nx = (const char *)memchr(c.p, '>', (size_t)(c.end - c.p));
            const char *stop = nx ? nx : c.end;
            if (take) {
                for (const char *q = c.p; q < stop; q++) seq_len += *q != '\n';
                rec.append(c.p, stop);
            }
            c.p = stop;
        } else {
            while (!c.eof() && *c.p != '>') {
                Line l = c.next();
                if (take && l.e > l.b) { rec.append(l.b, l.e); rec.push_back('\n'); seq_len += (size_t)(l.e - l.b); }
            }
        }
        if (take) {
            if (seq_len == 0) return;                // empty read = end of stream for the reference's loop
            out.put(rec.data(), rec.data() + rec.size());
            left--;
        }
        pos++;
    }
}

void extract_fastq(const std::string &t, bool gz, const BitVec &bv, uint64_t nb_reads, uint64_t left, Sink &out)
{
    Cursor c{t.data(), t.data() + t.size()};
    auto next_non_empty = [&](Line &l) -> bool {     // plain reader: getline, skipping empty lines
        while (!c.eof()) {
            l = c.next();
            if (l.e > l.b) return true;
        }
        return false;
    };
    for (uint64_t pos = 0; pos < nb_reads && left > 0; pos++) {
        Line l[4];
        bool ok = true;
        if (gz) {
            for (int i = 0; i < 4 && ok; i++) { ok = !c.eof(); if (ok) l[i] = c.next(); }
        } else {
            ok = next_non_empty(l[0]);
            if (ok) { ok = !c.eof(); if (ok) l[1] = c.next(); }     // the sequence line is taken as it comes
            if (ok) ok = next_non_empty(l[2]);
            if (ok) ok = next_non_empty(l[3]);
        }
        if (!ok) return;
        if (!bv.get(pos)) continue;
        if (l[2].e > l[2].b && *l[2].b != '+' && !gz) std::cerr << "Error\n";
        // gzip reader: the sequence is the line minus its LAST character (fastq_file.h:478-479), which is the
        // newline unless the file ends right after the sequence
        const char *se = l[1].e;
        if (gz && !l[1].nl && se > l[1].b) se--;
        if (se == l[1].b) return;                    // empty read: end of stream
        out.put(l[0].b, l[0].e); out.put('\n');
        out.put(l[1].b, se); out.put('\n');
        out.put(l[2].b, l[2].e); out.put('\n');
        out.put(l[3].b, l[3].e); out.put('\n');
        left--;
    }
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 3) print_usage();
    std::string input_file_name, bv_file_name, output_file_name;
    int arg_pos = 1;
    while (arg_pos < argc) {                                   // src/extract_reads.cpp:66-91
        std::string flag = argv[arg_pos];
        if (flag.empty() || flag[0] != '-') {
            if (input_file_name.empty()) input_file_name = flag;
            else if (bv_file_name.empty()) bv_file_name = flag;
            else std::cerr << "The mandatory files are already set, unknown file " << flag << " -> ignore\n";
        } else if (flag == "-o") {
            arg_pos++;
            if (arg_pos >= argc) {
                std::cerr << "Error, flag -o needs an argument\n";
                return 1;
            }
            output_file_name = argv[arg_pos];
        } else if (flag == "-h") {
            print_usage();
            return 0;
        } else if (flag == "-v") {
            std::cout << "\nextract_reads version " << version << "\n";
            return 0;
        } else {
            std::cerr << "Unknown option " << flag << "\n";
            print_usage();
            return 0;
        }
        arg_pos++;
    }
    if (input_file_name.empty()) {
        std::cerr << "Error: An input file name is needed -> exit\n";
        print_usage();
        return 0;
    } else if (bv_file_name.empty()) {
        std::cerr << "Error: A bv file name is needed -> exit\n";
        print_usage();
        return 0;
    }

    std::string text;
    Format fmt = Format::Unknown;
    bool gz = false;
    {
        // src/extract_reads.cpp:111-114 reports an unreadable file once itself ("file file" is its wording) before the
        // gzip probe reports it again and exits (:126-130)
        std::ifstream probe(input_file_name.c_str());
        if (!probe.good()) std::cerr << "Cannot open file file " << input_file_name << " -> ignore\n";
    }
    if (!load_text(input_file_name, text, fmt, gz, " -> ignore\n")) return 1;

    // the reference's record count = the size the vector must have (fasta_file.h:104-107, fastq_file.h:99-102)
    ParsedFile pf;
    if (fmt == Format::Fasta) parse_fasta(text, pf); else parse_fastq(text, pf);
    BitVec bv;
    bv.read(bv_file_name);
    if (bv.n != pf.nb_reads) {
        std::cerr << "Number of reads in " << input_file_name << " and boolean vector size are not equal -> quit\n";
        return 1;
    }
    uint64_t n_valid = 0;
    for (uint64_t i = 0; i < bv.n; i++) n_valid += bv.get(i);

    Sink *out = nullptr;
    FILE *fout = nullptr;
    gzFile gout = nullptr;
    if (gz) {                                                  // src/extract_reads.cpp:146-160
        if (output_file_name.empty()) {
            std::cerr << "Error, try to compress results but no output file name is given\n";
            return 1;
        }
        gout = gzopen(output_file_name.c_str(), "w6");
        if (!gout) {
            std::cerr << "Error, cannot open file " << output_file_name << "\n";
            return 1;
        }
        out = new GzSink(gout);
    } else if (!output_file_name.empty()) {
        fout = fopen(output_file_name.c_str(), "wb");
        if (!fout) {
            std::cerr << "Cannot write on file " << output_file_name << "\n";
            return 1;
        }
        out = new FileSink(fout);
    } else {
        out = new FileSink(stdout);
    }
    if (fmt == Format::Fasta) extract_fasta(text, gz, bv, pf.nb_reads, n_valid, *out);
    else extract_fastq(text, gz, bv, pf.nb_reads, n_valid, *out);
    out->flush_buf();
    delete out;
    if (gout) gzclose(gout);
    if (fout) fclose(fout);
    else fflush(stdout);
    return 0;
}
