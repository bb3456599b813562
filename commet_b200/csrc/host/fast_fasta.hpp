// Parallel loader for plain (uncompressed) FASTA files, for tools that keep whole read sets resident.
//
// Same record rules as parse_fasta in readers.hpp (= the reference's FastaFile, include/fasta_file.h:61-68 and
// :166-175): a record starts at every non-empty line whose first byte is '>', its sequence is the concatenation
// of the following non-empty lines up to the next '>' line, lines before the first '>' line are ignored, '\r'
// is data.  The file is mmap'ed and cut into chunks at record starts; every chunk is scanned twice by a pool of
// threads -- once to size it (records, sequence bytes), once to write its sequences and offsets straight into
// the caller's final buffers -- so there is no intermediate copy and no serial pass over the data.
// FASTQ and gzip files are not handled here: their record boundaries (non-empty lines / 4, fastq_file.h:60-67) and
// the inflate stream are sequential; callers fall back to parse_reads_file for them.
#pragma once
#include <fcntl.h>
#include <stdint.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <string>
#include <vector>

namespace commet_host {

struct FastaMap {
    std::string path;
    const char *data = nullptr;
    size_t size = 0;
    std::vector<size_t> cut;               // chunk c = [cut[c], cut[c+1]); every cut but the first is at a '>' line start
    std::vector<uint64_t> records, bytes;  // per chunk, filled by scan()
    uint64_t n_records = 0, n_bytes = 0;

    // true when the file is a plain FASTA this loader takes (first byte '>', file_manager.h:127-156)
    bool open(const std::string &fname, size_t chunk_bytes)
    {
        path = fname;
        int fd = ::open(fname.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat sb;
        if (fstat(fd, &sb) != 0 || sb.st_size == 0) { ::close(fd); return false; }
        size = (size_t)sb.st_size;
        void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (p == MAP_FAILED) return false;
        data = static_cast<const char *>(p);
        if (data[0] != '>') { close(); return false; }
        madvise(p, size, MADV_WILLNEED);
        cut.assign(1, 0);
        while (cut.back() + chunk_bytes < size) {
            // first line start at or after cut.back() + chunk_bytes whose line begins with '>'
            size_t pos = cut.back() + chunk_bytes - 1;
            while (pos < size) {
                const char *nl = static_cast<const char *>(memchr(data + pos, '\n', size - pos));
                if (!nl) { pos = size; break; }
                pos = (size_t)(nl - data) + 1;
                if (pos < size && data[pos] == '>') break;
            }
            if (pos >= size) break;
            cut.push_back(pos);
        }
        cut.push_back(size);
        records.assign(cut.size() - 1, 0);
        bytes.assign(cut.size() - 1, 0);
        return true;
    }

    void close()
    {
        if (data) munmap(const_cast<char *>(data), size);
        data = nullptr;
    }

    size_t n_chunks() const { return cut.size() - 1; }

    // one pass over chunk c.  WRITE = false: count records and sequence bytes.  WRITE = true: append the sequences
    // at seq + pos and store every record's start offset (pos at its header line) in offs[rec++].
    template <bool WRITE>
    void pass(size_t c, uint8_t *seq, uint64_t pos, uint64_t *offs, uint64_t rec)
    {
        const char *p = data + cut[c], *end = data + cut[c + 1];
        bool in_record = c > 0;                 // chunks after the first start at a header line
        uint64_t n_rec = 0, n_b = 0;
        while (p < end) {
            const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
            const char *le = nl ? nl : end;
            if (le > p) {
                if (*p == '>') {
                    in_record = true;
                    if (WRITE) offs[rec + n_rec] = pos + n_b;
                    n_rec++;
                } else if (in_record) {
                    if (WRITE) memcpy(seq + pos + n_b, p, (size_t)(le - p));
                    n_b += (uint64_t)(le - p);
                }
            }
            p = nl ? nl + 1 : end;
        }
        if (!WRITE) { records[c] = n_rec; bytes[c] = n_b; }
    }

    void finish_scan()
    {
        n_records = n_bytes = 0;
        for (size_t c = 0; c < n_chunks(); c++) { n_records += records[c]; n_bytes += bytes[c]; }
    }
};

}  // namespace commet_host
