// records.hpp -- the reference's only vector interchange format (src/utils/embeddings.rs:52-62):
// one record per line, `f32,f32,...,f32;payload`, floats as Rust's shortest round-trip decimal.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace vrod {

struct Record {
    std::vector<float> vec;
    std::string payload;
};

// Parses "f32,f32,...[;payload]".  Returns false (with *err) on an empty vector or a malformed number.
bool parse_record(const std::string &line, Record *out, std::string *err);
// Parses a comma-separated f32 list.
bool parse_vector(const std::string &text, std::vector<float> *out, std::string *err);
// Writes a record in the same format, floats with the shortest representation that round-trips.
std::string format_record(const std::vector<float> &vec, const std::string &payload);
// Reads a whole file of records; all vectors must have one dimension (embeddings.rs:35 takes
// embeddings[0].len() as THE dimension).  Blank lines are skipped.
bool read_records_file(const std::string &path, std::vector<float> *rows, std::vector<std::string> *payloads,
                       uint32_t *dim, std::string *err);

}  // namespace vrod
