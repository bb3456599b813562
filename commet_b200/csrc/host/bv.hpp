// .bv boolean-vector files, format of include/boolean_vector.h:302-414:
//   <comment> "\n#" <n_bits> "\n" <payload: n_bits/8+1 bytes, LSB first>
#pragma once
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace commet_host {

struct BitVec {
    std::string comment;
    uint64_t n = 0;                 // bits
    std::vector<uint8_t> bytes;     // n/8+1

    static uint64_t payload_bytes(uint64_t n) { return n / 8 + 1; }

    void init_false(uint64_t size) { n = size; bytes.assign(payload_bytes(size), 0); }
    void init_true(uint64_t size)   // boolean_vector.h:138-153: padding bits cleared
    {
        n = size;
        bytes.assign(payload_bytes(size), 0xFF);
        for (uint64_t i = n; i < bytes.size() * 8; i++) bytes[i / 8] &= (uint8_t)~(1u << (i % 8));
    }
    bool get(uint64_t i) const { return (bytes[i / 8] >> (i % 8)) & 1u; }
    void set(uint64_t i) { bytes[i / 8] |= (uint8_t)(1u << (i % 8)); }
    void unset(uint64_t i) { bytes[i / 8] &= (uint8_t)~(1u << (i % 8)); }

    // boolean_vector.h:347-414
    void read(const std::string &file_name)
    {
        int fd = open(file_name.c_str(), O_RDONLY);
        if (fd == -1) {
            std::cerr << "Error opening file " << file_name << " -> exit\n";
            exit(1);
        }
        struct stat sb;
        if (fstat(fd, &sb) == -1) {
            std::cerr << "Error getting statistics from file " << file_name << " -> exit\n";
            close(fd);
            exit(1);
        }
        std::string raw((size_t)sb.st_size, '\0');
        size_t got = 0;
        while (got < raw.size()) {
            ssize_t r = ::read(fd, &raw[got], raw.size() - got);
            if (r <= 0) break;
            got += (size_t)r;
        }
        close(fd);
        raw.resize(got);
        size_t i = 0;
        comment.clear();
        while (i < raw.size() && raw[i] != '#') comment += raw[i++];
        i++;
        if (!comment.empty()) comment.erase(comment.size() - 1);     // the '\n' before '#'
        std::string num;
        while (i < raw.size() && raw[i] != '\n') num += raw[i++];
        i++;
        if (num.empty()) {
            std::cerr << "Error, boolean vector does not contain its size\n";
            exit(1);
        }
        init_false((uint64_t)(unsigned long)atoi(num.c_str()));
        size_t avail = i < raw.size() ? raw.size() - i : 0;
        memcpy(bytes.data(), raw.data() + std::min(i, raw.size()), std::min(avail, bytes.size()));
    }

    std::string header() const
    {
        std::stringstream s;
        s << comment << "\n#" << n << "\n";
        return s.str();
    }

    // boolean_vector.h:302-346 (mode 0600, truncate)
    void write(const std::string &file_name) const
    {
        std::string h = header();
        int fd = open(file_name.c_str(), O_RDWR | O_CREAT | O_TRUNC, (mode_t)0600);
        if (fd == -1) {
            std::cerr << "Error opening file " << file_name << " -> exit\n";
            exit(1);
        }
        std::string all = h;
        all.append(reinterpret_cast<const char *>(bytes.data()), bytes.size());
        size_t done = 0;
        while (done < all.size()) {
            ssize_t w = ::write(fd, all.data() + done, all.size() - done);
            if (w <= 0) {
                std::cerr << "Error writing last byte of " << file_name << " -> exit\n";
                close(fd);
                exit(1);
            }
            done += (size_t)w;
        }
        close(fd);
    }

    // boolean_vector.h:287-295: header + raw bytes on stdout
    void print_stdout() const
    {
        std::cout << comment << "\n#" << n << "\n";
        std::cout.write(reinterpret_cast<const char *>(bytes.data()), (std::streamsize)bytes.size());
    }
};

}  // namespace commet_host
