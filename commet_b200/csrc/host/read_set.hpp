// Host-side equivalent of FileManager (include/file_manager.h:39-315): a set
// of read files seen as one stream of VALID reads (those whose bit is set in
// the file's input boolean vector), per-file output vectors, and the naming
// of the .bv files written by save_bv (:245-252).
#pragma once
#include "bv.hpp"
#include "readers.hpp"

#include <string>
#include <utility>
#include <vector>

namespace commet_host {

struct SetFile {
    std::string fname;
    uint64_t nb_reads = 0;
    BitVec in_bv;                       // which reads take part   (ReadFile::bv)
    BitVec out_bv;                      // result of the search    (FileManager::file_bvs)
    std::vector<uint64_t> valid_pos;    // record index of every valid read, in stream order
    ParsedFile data;                    // sequences (released once the stream is built unless kept)
};

struct ReadSet {
    std::string nickname;
    std::vector<SetFile> files;
    ByteVec bases;                      // valid-read stream
    std::vector<uint64_t> offs{0};

    uint64_t n_valid() const { return offs.size() - 1; }

    // FileManager::addFile (file_manager.h:117-222).  bv_name empty -> all reads valid.
    // Returns false if the file was ignored (message already printed).
    // `parsed`: the file's records if they were parsed ahead of time (pass.hpp: ParseAhead), else null
    bool add_file(const std::string &fname, const std::string &bv_name, ParsedFile *parsed = nullptr)
    {
        SetFile sf;
        // file_manager.h:119-143: an unreadable file is reported by the first-byte probe ("file file" is the reference's
        // wording), then again by the gzip probe, which exits; only a file of unknown format is ignored
        bool readable;
        {
            std::ifstream probe(fname.c_str());
            readable = probe.good();
        }
        if (!readable && !bv_name.empty()) {
            // the two-argument addFile (file_manager.h:167-175, the form Commet.py's "file,bv" entries take): one
            // message, the file is skipped and the run goes on
            std::cerr << "Cannot open file " << fname << " -> ignore\n";
            return false;
        }
        if (!readable) std::cerr << "Cannot open file file " << fname << " -> ignore\n";
        if (parsed && readable) {
            sf.data = std::move(*parsed);
        } else if (!parse_reads_file(fname, sf.data, " -> ignore\n")) {
            if (!readable) exit(1);
            return false;
        }
        sf.fname = fname;
        sf.nb_reads = sf.data.nb_reads;
        if (bv_name.empty()) {
            sf.in_bv.init_true(sf.nb_reads);
        } else {
            sf.in_bv.read(bv_name);
            if (sf.in_bv.n != sf.nb_reads) {       // fasta_file.h:104-107
                std::cerr << "Number of reads in " << fname << " and boolean vector size are not equal -> quit\n";
                exit(1);
            }
        }
        sf.out_bv.init_false(sf.nb_reads);
        files.push_back(std::move(sf));
        return true;
    }

    // (Re)build the valid-read stream from the files' in_bv.  keep=false releases the parsed
    // sequences afterwards (single pass tools); keep=true is for -f, which re-streams the set.
    void build_stream(bool keep)
    {
        bases.clear();
        offs.assign(1, 0);
        const bool single_all = files.size() == 1 && !keep;
        for (SetFile &f : files) {
            f.valid_pos.clear();
            ParsedFile &pf = f.data;
            bool all_valid = true;
            for (uint64_t r = 0; r < f.nb_reads && all_valid; r++) all_valid = f.in_bv.get(r);
            for (uint64_t r = 0; r < f.nb_reads; r++) {
                if (!f.in_bv.get(r)) continue;
                if (pf.off[r + 1] == pf.off[r]) {
                    // the reference treats an empty read as end-of-set and then runs past its file
                    // table (file_manager.h:88-97): undefined behaviour there, a clean error here
                    std::cerr << "Error: record " << r << " of " << f.fname << " has an empty sequence -> exit\n";
                    exit(1);
                }
                f.valid_pos.push_back(r);
            }
            if (single_all && all_valid) {         // the parsed buffer IS the stream
                bases = std::move(pf.seq);
                offs = std::move(pf.off);
            } else {
                for (uint64_t r : f.valid_pos) {
                    bases.insert(bases.end(), pf.seq.begin() + (std::ptrdiff_t)pf.off[r],
                                 pf.seq.begin() + (std::ptrdiff_t)pf.off[r + 1]);
                    offs.push_back(bases.size());
                }
            }
            if (!keep) {
                ByteVec().swap(pf.seq);
                std::vector<uint64_t>().swap(pf.off);
            }
        }
    }

    void clear_out()
    {
        for (SetFile &f : files) f.out_bv.init_false(f.nb_reads);
    }

    // scatter a stream-ordered tag vector (.bv payload layout) into the per-file output vectors
    void scatter_tags(const std::vector<uint8_t> &tags)
    {
        uint64_t s = 0;
        for (SetFile &f : files)
            for (uint64_t pos : f.valid_pos) {
                if ((tags[s / 8] >> (s % 8)) & 1u) f.out_bv.set(pos);
                s++;
            }
    }

    // apply_bv_on_files (file_manager.h:277-285): the tagged reads become the valid ones
    void apply_out_as_input()
    {
        for (SetFile &f : files) {
            std::string c = f.in_bv.comment;
            f.in_bv = f.out_bv;
            f.in_bv.comment = c;
            f.out_bv.init_false(f.nb_reads);
        }
    }

    // FileManager::save_bv (file_manager.h:245-252)
    void save_bv(const std::string &directory, const std::string &suffix)
    {
        for (SetFile &f : files) {
            std::string base = f.fname.substr(f.fname.rfind("/") + 1);
            f.out_bv.comment = f.fname + " in " + suffix;
            f.out_bv.write(directory + "/" + base + "_in_" + suffix + ".bv");
        }
    }
};

}  // namespace commet_host
