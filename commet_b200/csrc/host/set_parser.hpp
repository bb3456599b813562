// File-of-files grammar of include/set_parser.h:46-102:
//   name:file[,bv][;file[,bv]]*      one set per non-empty line
// Spaces are trimmed around files and bvs but NOT around the name; a line
// without ':' is named SET<n>; sets come back in std::map (sorted-name) order.
#pragma once
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace commet_host {

inline void trim_spaces(std::string &s)
{
    size_t b = 0, e = s.size();
    while (b < e && s[b] == ' ') b++;
    while (e > b && s[e - 1] == ' ') e--;
    s = s.substr(b, e - b);
}

struct SetSpec {
    std::vector<std::string> files, bvs;
};

inline std::map<std::string, SetSpec> read_sets(const std::string &file_name)
{
    std::map<std::string, SetSpec> sets;
    std::ifstream in(file_name.c_str());
    if (!in.good()) {
        std::cerr << "Cannot read file " << file_name << "\n";
        exit(1);
    }
    int nb_sets = 0;
    std::string line;
    while (in.good()) {
        line.clear();
        std::getline(in, line);
        if (line.empty()) continue;
        nb_sets++;
        std::string tag;
        size_t colon = line.find(':');
        if (colon != std::string::npos) {
            tag = line.substr(0, colon);
            line = line.substr(colon + 1);
        } else {
            std::stringstream t;
            t << "SET" << nb_sets;
            tag = t.str();
        }
        SetSpec spec;
        auto push = [&spec](std::string item) {
            trim_spaces(item);
            std::string bv;
            size_t comma = item.find(',');
            if (comma != std::string::npos) {
                bv = item.substr(comma + 1);
                trim_spaces(bv);
                item = item.substr(0, comma);
                trim_spaces(item);
            }
            spec.files.push_back(item);
            spec.bvs.push_back(bv);
        };
        size_t semi;
        while (!line.empty() && (semi = line.find(';')) != std::string::npos) {
            push(line.substr(0, semi));
            line = line.substr(semi + 1);
        }
        push(line);
        sets[tag] = spec;
    }
    return sets;
}

}  // namespace commet_host
