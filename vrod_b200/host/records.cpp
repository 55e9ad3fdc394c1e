#include "records.hpp"

#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>

namespace vrod {

bool parse_vector(const std::string &text, std::vector<float> *out, std::string *err) {
    out->clear();
    size_t pos = 0;
    while (pos <= text.size()) {
        size_t end = text.find(',', pos);
        if (end == std::string::npos) end = text.size();
        size_t a = pos, b = end;
        while (a < b && (text[a] == ' ' || text[a] == '\t')) ++a;
        while (b > a && (text[b - 1] == ' ' || text[b - 1] == '\t' || text[b - 1] == '\r')) --b;
        if (a == b) {
            if (err) *err = "empty vector component";
            return false;
        }
        const std::string tok = text.substr(a, b - a);
        char *stop = nullptr;
        const float v = std::strtof(tok.c_str(), &stop);
        if (stop == tok.c_str() || *stop != 0) {
            if (err) *err = "malformed number '" + tok + "'";
            return false;
        }
        out->push_back(v);
        pos = end + 1;
        if (end == text.size()) break;
    }
    if (out->empty()) {
        if (err) *err = "empty vector";
        return false;
    }
    return true;
}

bool parse_record(const std::string &line, Record *out, std::string *err) {
    const size_t semi = line.find(';');
    const std::string vec = semi == std::string::npos ? line : line.substr(0, semi);
    out->payload = semi == std::string::npos ? std::string() : line.substr(semi + 1);
    while (!out->payload.empty() && (out->payload.back() == '\n' || out->payload.back() == '\r')) out->payload.pop_back();
    return parse_vector(vec, &out->vec, err);
}

std::string format_record(const std::vector<float> &vec, const std::string &payload) {
    std::string s;
    char buf[64];
    for (size_t i = 0; i < vec.size(); ++i) {
        if (i) s += ',';
        auto r = std::to_chars(buf, buf + sizeof(buf), vec[i]);  // shortest round-trip, like Rust's f32::to_string
        s.append(buf, r.ptr);
    }
    s += ';';
    s += payload;
    return s;
}

bool read_records_file(const std::string &path, std::vector<float> *rows, std::vector<std::string> *payloads,
                       uint32_t *dim, std::string *err) {
    std::ifstream f(path);
    if (!f) {
        if (err) *err = "cannot open '" + path + "'";
        return false;
    }
    rows->clear();
    payloads->clear();
    *dim = 0;
    std::string line;
    Record rec;
    size_t lineno = 0;
    while (std::getline(f, line)) {
        ++lineno;
        if (line.empty() || line == "\r") continue;
        std::string perr;
        if (!parse_record(line, &rec, &perr)) {
            if (err) *err = path + ":" + std::to_string(lineno) + ": " + perr;
            return false;
        }
        if (*dim == 0) *dim = (uint32_t)rec.vec.size();
        if (rec.vec.size() != *dim) {
            if (err) *err = path + ":" + std::to_string(lineno) + ": dimension " + std::to_string(rec.vec.size()) +
                            " differs from " + std::to_string(*dim);
            return false;
        }
        rows->insert(rows->end(), rec.vec.begin(), rec.vec.end());
        payloads->push_back(rec.payload);
    }
    if (payloads->empty()) {
        if (err) *err = "'" + path + "' holds no records";
        return false;
    }
    return true;
}

}  // namespace vrod
