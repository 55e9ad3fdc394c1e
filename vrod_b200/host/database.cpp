#include "database.hpp"

#include <fcntl.h>
#include <unistd.h>

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

Database Database::load(const std::filesystem::path &dir, std::vector<int> devices) {
    namespace fs = std::filesystem;
    std::ifstream cfg(dir / "vr_config");
    if (!cfg) throw std::runtime_error("'" + dir.string() + "' is not a vRod database directory (no vr_config)");
    Database db(std::move(devices));
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

static void flush_to_disk(const std::filesystem::path &p) {
    const int fd = ::open(p.c_str(), O_RDONLY);
    if (fd >= 0) {
        ::fsync(fd);
        ::close(fd);
    }
}

void Database::save() {
    namespace fs = std::filesystem;
    if (path.empty()) return;
    std::vector<fs::path> written;   // final names; each was written as <name>.tmp
    auto tmp_of = [](const fs::path &p) { return fs::path(p.string() + ".tmp"); };
    try {
        const fs::path cfg_path = path / "vr_config";
        std::ofstream cfg(tmp_of(cfg_path), std::ios::trunc);
        if (!cfg) throw std::runtime_error("cannot write '" + tmp_of(cfg_path).string() + "'");
        written.push_back(cfg_path);
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
            const fs::path rows_path = path / (name + ".vrc"), pay_path = path / (name + ".payloads");
            written.push_back(rows_path);
            if (vrod_collection_save(c, tmp_of(rows_path).string().c_str()) != VROD_OK)
                throw std::runtime_error(std::string("saving collection '") + name + "': " + vrod_last_error());
            written.push_back(pay_path);
            std::ofstream pf(tmp_of(pay_path), std::ios::trunc);
            for (const std::string &p : payloads[name]) pf << p << '\n';
            pf.flush();
            if (!pf) throw std::runtime_error("short write to '" + tmp_of(pay_path).string() + "'");
            cfg << "collection " << name << '\n';
        }
        for (const auto &kv : pending)
            cfg << "pending " << kv.first << ' ' << (kv.second.metric == VROD_COSINE ? 1 : 0) << ' ' << kv.second.capacity << '\n';
        cfg.flush();
        if (!cfg) throw std::runtime_error("short write to '" + tmp_of(cfg_path).string() + "'");
        cfg.close();
        for (const fs::path &p : written) flush_to_disk(tmp_of(p));
        // rows and payloads first, the index (vr_config) last
        for (size_t i = written.size(); i-- > 0;) fs::rename(tmp_of(written[i]), written[i]);
    } catch (...) {
        std::error_code ec;
        for (const fs::path &p : written) fs::remove(tmp_of(p), ec);
        throw;
    }
    dirty = false;
}

Database::~Database() {
    if (ctx_) vrod_ctx_destroy(ctx_);
}

Database::Database(Database &&o) noexcept
    : dirty(o.dirty), pending(std::move(o.pending)), payloads(std::move(o.payloads)), last(std::move(o.last)),
      path(std::move(o.path)), devices_(std::move(o.devices_)), ctx_(o.ctx_) {
    o.ctx_ = nullptr;
}

vrod_ctx *Database::ctx() {
    if (!ctx_) {
        if (vrod_ctx_create_multi(devices_.data(), (int)devices_.size(), &ctx_) != VROD_OK) throw std::runtime_error(vrod_last_error());
    }
    return ctx_;
}

}  // namespace vrod
