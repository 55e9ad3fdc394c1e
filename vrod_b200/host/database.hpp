// database.hpp -- host-side mirror of the reference's Database (src/database/mod.rs:6-22), C++17.
//
// The reference struct holds only a path; its collections are a TODO (mod.rs:8).  Here the Database
// owns the C-ABI context (include/vrod_knn.h) whose collections live in GPU memory, plus the text
// payload of every record (the word after ';' in the reference's record format,
// src/utils/embeddings.rs:52-62), which search results print next to the id.
#pragma once
#include <cstdint>
#include <filesystem>
#include <map>
#include <optional>
#include <string>
#include <vector>

#include "../../include/vrod_knn.h"

namespace vrod {

// What a command leaves behind: `trait Command::execute(&self)` returns nothing and has no error
// channel (src/command/types.rs:5-7), so results and failures leave through the Database.
struct CommandResult {
    bool ok = true;
    std::string error;                 // empty when ok
    std::vector<uint64_t> ids;         // SEARCH: k ids
    std::vector<float> dist;           // SEARCH: k distances
    std::vector<std::string> payload;  // SEARCH: payload of each hit ("" if none)
    std::vector<std::string> names;    // LISTCOLLECTIONS
    uint64_t first_id = 0, inserted = 0;
};

struct CollectionSpec {   // CREATE before the first INSERT fixes the dimension
    std::optional<uint32_t> dim;
    vrod_metric metric = VROD_EUCLIDEAN;
    uint64_t capacity = 1u << 20;
};

class Database {
  public:
    // Database::new(path, name) (mod.rs:13-17): creates path/name with empty vr_config and vr_wal,
    // io error AlreadyExists if the directory exists (src/database/setup.rs:3-26).
    static Database create(const std::filesystem::path &path, const std::string &name);
    // An in-memory database on one GPU, or row-sharded over several GPUs driven by this one process
    // (vrod_ctx_create_multi: the reference's caller is a single-threaded process, src/main.rs:42).
    explicit Database(int device = 0) : devices_{device} {}
    explicit Database(std::vector<int> devices) : devices_(std::move(devices)) {}
    // Database::load(path) (mod.rs:19-21, todo!() upstream): opens a directory made by create() and loads
    // every collection listed in its vr_config into GPU memory.  Throws std::runtime_error on a directory
    // without vr_config or a damaged collection file.
    static Database load(const std::filesystem::path &dir, std::vector<int> devices = {0});
    // Writes vr_config, one <name>.vrc rows file (vrod_collection_save) and one <name>.payloads text file per
    // collection into `path`: everything goes to <file>.tmp first, is flushed to disk, and is renamed into place
    // only after every collection has been written, so a failure midway (full disk) leaves the previous state.
    // vr_wal is left untouched: there is no write-ahead log in this build.
    void save();
    bool dirty = false;   // set by the commands that change collections
    ~Database();
    Database(Database &&o) noexcept;
    Database(const Database &) = delete;

    vrod_ctx *ctx();  // created on first use; throws std::runtime_error without a GPU (no CPU path)

    std::map<std::string, CollectionSpec> pending;                 // created, not yet materialised
    std::map<std::string, std::vector<std::string>> payloads;      // collection -> payload by id
    CommandResult last;                                            // result of the last command
    std::filesystem::path path;

  private:
    std::vector<int> devices_{0};
    vrod_ctx *ctx_ = nullptr;
};

}  // namespace vrod
