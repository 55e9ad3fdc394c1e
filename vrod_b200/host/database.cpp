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

Database::~Database() {
    if (ctx_) vrod_ctx_destroy(ctx_);
}

Database::Database(Database &&o) noexcept
    : pending(std::move(o.pending)), payloads(std::move(o.payloads)), last(std::move(o.last)), path(std::move(o.path)),
      device_(o.device_), ctx_(o.ctx_) {
    o.ctx_ = nullptr;
}

vrod_ctx *Database::ctx() {
    if (!ctx_) {
        if (vrod_ctx_create(device_, &ctx_) != VROD_OK) throw std::runtime_error(vrod_last_error());
    }
    return ctx_;
}

}  // namespace vrod
