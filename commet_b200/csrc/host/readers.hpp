// FASTA / FASTQ (plain or gzip) readers producing, in ONE pass, what the
// reference obtains with a counting pass plus a reading pass:
//   - the record count            (fasta_file.h:61-68: lines starting with '>';
//                                  fastq_file.h:60-67: non-empty lines / 4)
//   - every record's sequence     (fasta_file.h:166-175: non-empty lines up to the next '>' line, concatenated;
//                                  fastq_file.h:132-180: the line after the '@' line)
// Sequences are appended to one contiguous byte buffer (`seq`) with `off`
// giving record boundaries, ready to be handed to commet_reads_upload.
#pragma once
#include <stdint.h>
#include <zlib.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

#include "fast_fasta.hpp"

namespace commet_host {

enum class Format { Fasta, Fastq, Unknown };

// byte buffer whose resize() does not zero-fill: the parallel loader sizes the sequence buffer first and writes every
// byte of it afterwards -- value-initialising a gigabyte on one core cost as much as parsing it on sixteen
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAlloc<U>; };
    using std::allocator<T>::allocator;
    template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
using ByteVec = std::vector<uint8_t, DefaultInitAlloc<uint8_t>>;

struct ParsedFile {
    std::string fname;
    Format format = Format::Unknown;
    bool gz = false;
    uint64_t nb_reads = 0;            // the reference's count (bit-vector size)
    std::vector<uint64_t> off;        // nb_reads+1 offsets into seq
    ByteVec seq;
};

inline bool slurp_plain(const std::string &fname, std::string &out)
{
    std::ifstream f(fname.c_str(), std::ios::binary);
    if (!f.good()) return false;
    f.seekg(0, std::ios::end);
    std::streamoff sz = f.tellg();
    f.seekg(0);
    out.resize((size_t)sz);
    if (sz) f.read(&out[0], sz);
    return true;
}

inline bool slurp_gz(const std::string &fname, std::string &out)
{
    gzFile g = gzopen(fname.c_str(), "r");
    if (!g) return false;
    gzbuffer(g, 1 << 20);
    out.clear();
    std::vector<char> buf(1 << 22);
    int n;
    while ((n = gzread(g, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), (size_t)n);
    gzclose(g);
    return true;
}

// Sniffing of file_manager.h:117-156 / filter_reads.cpp:121-154: first byte '>' or '@', else gzip.
// Returns false (with the reference's message on stderr) when the file cannot be used.
inline bool load_text(const std::string &fname, std::string &text, Format &fmt, bool &gz, const char *who)
{
    std::ifstream probe(fname.c_str());
    if (!probe.good()) {
        std::cerr << "Cannot open file " << fname << who;
        return false;
    }
    int c = probe.get();
    probe.close();
    gz = false;
    if (c == '>' || c == '@') {
        if (!slurp_plain(fname, text)) return false;
    } else {
        if (!slurp_gz(fname, text)) {
            std::cerr << "Cannot open file " << fname << who;
            return false;
        }
        gz = true;
        c = text.empty() ? -1 : (unsigned char)text[0];
    }
    fmt = c == '>' ? Format::Fasta : c == '@' ? Format::Fastq : Format::Unknown;
    if (fmt == Format::Unknown) {
        std::cerr << "Unknown format: " << fname << who;
        return false;
    }
    return true;
}

inline void parse_fasta(const std::string &t, ParsedFile &pf)
{
    const char *p = t.data(), *end = p + t.size();
    pf.seq.reserve(t.size());
    bool in_record = false;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        if (le > p) {
            if (*p == '>') {
                pf.off.push_back(pf.seq.size());
                in_record = true;
            } else if (in_record) {
                pf.seq.insert(pf.seq.end(), (const uint8_t *)p, (const uint8_t *)le);
            }
        }
        p = nl ? nl + 1 : end;
    }
    pf.nb_reads = pf.off.size();
    pf.off.push_back(pf.seq.size());
}

inline void parse_fastq(const std::string &t, ParsedFile &pf)
{
    const char *p = t.data(), *end = p + t.size();
    pf.seq.reserve(t.size() / 2);
    // counting rule first: non-empty lines / 4
    uint64_t non_empty = 0;
    for (const char *q = p; q < end;) {
        const char *nl = (const char *)memchr(q, '\n', (size_t)(end - q));
        const char *le = nl ? nl : end;
        if (le > q) non_empty++;
        q = nl ? nl + 1 : end;
    }
    pf.nb_reads = non_empty / 4;
    auto next_line = [&](const char *&b, const char *&e) -> bool {
        if (p >= end) return false;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        b = p;
        e = nl ? nl : end;
        p = nl ? nl + 1 : end;
        return true;
    };
    auto next_non_empty = [&](const char *&b, const char *&e) -> bool {
        while (next_line(b, e))
            if (e > b) return true;
        return false;
    };
    for (uint64_t r = 0; r < pf.nb_reads; r++) {
        const char *b, *e;
        pf.off.push_back(pf.seq.size());
        if (!next_non_empty(b, e)) continue;            // '@' line
        if (!next_line(b, e)) continue;                 // sequence line, taken as is (may be empty)
        pf.seq.insert(pf.seq.end(), (const uint8_t *)b, (const uint8_t *)e);
        if (!next_non_empty(b, e)) continue;            // '+' line
        if (b[0] != '+') std::cerr << "Error\n";        // fastq_file.h:158-160
        next_non_empty(b, e);                           // quality line
    }
    pf.off.push_back(pf.seq.size());
}

// Plain FASTA files of at least COMMET_B200_FAST_FASTA_MIN bytes (default 64 MiB): mmap + chunked two-pass loader
// on up to 16 threads (fast_fasta.hpp; same record rules as parse_fasta, checked against it in
// tests/test_nxn_host.py and tests/test_host_cpu.py) instead of slurping the file and parsing it on one core.
inline bool parse_fasta_parallel(const std::string &fname, ParsedFile &pf)
{
    uint64_t min_bytes = 64ull << 20;
    if (const char *e = getenv("COMMET_B200_FAST_FASTA_MIN")) min_bytes = strtoull(e, nullptr, 10);
    FastaMap m;
    if (!m.open(fname, 16u << 20)) return false;
    if (m.size < min_bytes) { m.close(); return false; }
    const size_t nc = m.n_chunks();
    const unsigned nt = (unsigned)std::min<size_t>(std::max(1u, std::min(std::thread::hardware_concurrency(), 16u)), nc);
    auto parallel_for = [&](auto fn) {
        std::atomic<size_t> next{0};
        std::vector<std::thread> th;
        for (unsigned w = 1; w < nt; w++)
            th.emplace_back([&]() { for (size_t c; (c = next.fetch_add(1)) < nc;) fn(c); });
        for (size_t c; (c = next.fetch_add(1)) < nc;) fn(c);
        for (auto &t : th) t.join();
    };
    parallel_for([&](size_t c) { m.pass<false>(c, nullptr, 0, nullptr, 0); });
    m.finish_scan();
    pf.seq.resize(m.n_bytes);
    pf.off.resize(m.n_records + 1);
    pf.off[m.n_records] = m.n_bytes;
    std::vector<uint64_t> pos(nc + 1, 0), rec(nc + 1, 0);
    for (size_t c = 0; c < nc; c++) { pos[c + 1] = pos[c] + m.bytes[c]; rec[c + 1] = rec[c] + m.records[c]; }
    parallel_for([&](size_t c) { m.pass<true>(c, pf.seq.data(), pos[c], pf.off.data(), rec[c]); });
    pf.nb_reads = m.n_records;
    pf.format = Format::Fasta;
    pf.gz = false;
    m.close();
    return true;
}

inline bool parse_reads_file(const std::string &fname, ParsedFile &pf, const char *who)
{
    std::string text;
    pf.fname = fname;
    if (parse_fasta_parallel(fname, pf)) return true;
    if (!load_text(fname, text, pf.format, pf.gz, who)) return false;
    if (pf.format == Format::Fasta) parse_fasta(text, pf);
    else parse_fastq(text, pf);
    return true;
}

}  // namespace commet_host
