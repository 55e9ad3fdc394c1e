#include "database.hpp"

#include <fstream>
#include <stdexcept>

namespace vrod {

Database Database::create(const std::filesystem::path &path, const std::string &name) {
    namespace fs = std::filesystem;
    const fs::path dir = path / name;
    if (fs::exists(dir))
        throw fs::filesystem_error("Directory with the name '" + name + "' already exists in '" + path.string() + "'",
                                   std::make_error_code(std::errc::file_exists));
    fs::create_directory(dir);               // not create_directories: the reference uses fs::create_dir
    std::ofstream(dir / "vr_config").flush();
    std::ofstream(dir / "vr_wal").flush();
    Database db;
    db.path = path;
    return db;
}

Database Database::load(const std::filesystem::path &dir, int device) {
    namespace fs = std::filesystem;
    std::ifstream cfg(dir / "vr_config");
    if (!cfg) throw std::runtime_error("'" + dir.string() + "' is not a vRod database directory (no vr_config)");
    Database db(device);
    db.path = dir;
    std::string kind, name;
    while (cfg >> kind >> name) {
        if (kind == "collection") {
            vrod_collection *c = nullptr;
            if (vrod_collection_load(db.ctx(), name.c_str(), (dir / (name + ".vrc")).string().c_str(), 0, &c) != VROD_OK)
                throw std::runtime_error(std::string("loading collection '") + name + "': " + vrod_last_error());
            auto &pl = db.payloads[name];
            std::ifstream pf(dir / (name + ".payloads"));
            std::string line;
            while (std::getline(pf, line)) pl.push_back(line);
        } else if (kind == "pending") {
            CollectionSpec spec;
            int metric = 0;
            cfg >> metric >> spec.capacity;
            spec.metric = metric ? VROD_COSINE : VROD_EUCLIDEAN;
            db.pending[name] = spec;
            db.payloads[name];
        } else {
            throw std::runtime_error("vr_config: unknown entry '" + kind + "'");
        }
    }
    return db;
}

void Database::save() {
    namespace fs = std::filesystem;
    if (path.empty()) return;
    std::ofstream cfg(path / "vr_config", std::ios::trunc);
    size_t need = 0;
    vrod_collection_list(ctx(), nullptr, 0, &need);
    std::string names(need, '\0');
    vrod_collection_list(ctx(), names.data(), need, &need);
    names.resize(need ? need - 1 : 0);
    size_t pos = 0;
    while (pos < names.size()) {
        size_t e = names.find('\n', pos);
        if (e == std::string::npos) e = names.size();
        const std::string name = names.substr(pos, e - pos);
        pos = e + 1;
        vrod_collection *c = nullptr;
        if (vrod_collection_get(ctx(), name.c_str(), &c) != VROD_OK) continue;
        if (vrod_collection_save(c, (path / (name + ".vrc")).string().c_str()) != VROD_OK)
            throw std::runtime_error(std::string("saving collection '") + name + "': " + vrod_last_error());
        std::ofstream pf(path / (name + ".payloads"), std::ios::trunc);
        for (const std::string &p : payloads[name]) pf << p << '\n';
        cfg << "collection " << name << '\n';
    }
    for (const auto &kv : pending)
        cfg << "pending " << kv.first << ' ' << (kv.second.metric == VROD_COSINE ? 1 : 0) << ' ' << kv.second.capacity << '\n';
    dirty = false;
}

Database::~Database() {
    if (ctx_) vrod_ctx_destroy(ctx_);
}

Database::Database(Database &&o) noexcept
    : dirty(o.dirty), pending(std::move(o.pending)), payloads(std::move(o.payloads)), last(std::move(o.last)),
      path(std::move(o.path)), device_(o.device_), ctx_(o.ctx_) {
    o.ctx_ = nullptr;
}

vrod_ctx *Database::ctx() {
    if (!ctx_) {
        if (vrod_ctx_create(device_, &ctx_) != VROD_OK) throw std::runtime_error(vrod_last_error());
    }
    return ctx_;
}

}  // namespace vrod
