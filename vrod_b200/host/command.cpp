#include "command.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "records.hpp"

namespace vrod {

namespace {

std::string opt_repr(const OptStr &s) { return s ? "Some(\"" + *s + "\")" : "None"; }

void set_error(const Db &db, const std::string &msg) {
    db->last = CommandResult{};
    db->last.ok = false;
    db->last.error = msg;
    std::fprintf(stderr, "error: %s\n", msg.c_str());
}

bool api(const Db &db, vrod_status st) {
    if (st == VROD_OK) return true;
    set_error(db, vrod_last_error());
    return false;
}

std::vector<std::string> split(const std::string &s, char sep) {
    std::vector<std::string> out;
    size_t pos = 0;
    while (true) {
        const size_t e = s.find(sep, pos);
        out.push_back(s.substr(pos, e == std::string::npos ? std::string::npos : e - pos));
        if (e == std::string::npos) break;
        pos = e + 1;
    }
    return out;
}

std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}

// Returns the library collection, materialising a pending CREATE now that the dimension is known.
vrod_collection *resolve(const Db &db, const std::string &name, std::optional<uint32_t> dim_hint, uint64_t min_capacity) {
    vrod_collection *c = nullptr;
    if (vrod_collection_get(db->ctx(), name.c_str(), &c) == VROD_OK) return c;
    auto it = db->pending.find(name);
    if (it == db->pending.end()) {
        set_error(db, "no collection '" + name + "'");
        return nullptr;
    }
    CollectionSpec spec = it->second;
    if (!spec.dim) spec.dim = dim_hint;
    if (!spec.dim) {
        set_error(db, "collection '" + name + "' has no dimension yet (insert a record first)");
        return nullptr;
    }
    const uint64_t cap = std::max<uint64_t>(spec.capacity, min_capacity);
    if (!api(db, vrod_collection_create(db->ctx(), name.c_str(), *spec.dim, spec.metric, cap, &c))) return nullptr;
    db->pending.erase(it);
    return c;
}

void unsupported(const Db &db, const char *what) {
    set_error(db, std::string(what) + " is outside the SEARCH hot path this build implements (DESIGN.md section 7)");
}

}  // namespace

std::string Command::describe() const {
    return std::string(type_name()) + "{collection_name=" + opt_repr(collection_name) + ",arg=" + opt_repr(arg) + "}";
}

// CREATE -a name[;dim[;metric[;capacity]]]
void CreateCollectionCommand::execute() const {
    if (!collection_name || collection_name->empty()) return set_error(db, "CREATE needs a collection name in the argument");
    const std::vector<std::string> f = split(*collection_name, ';');
    const std::string &name = f[0];
    {   // the library's rule (vrod_collection_create), applied here too because a CREATE without a dimension only
        // materialises at the first INSERT: names become file names and vr_config tokens
        bool ok = !name.empty() && name.size() <= 200 && name != "." && name != "..";
        for (const char ch : name) ok = ok && (std::isalnum((unsigned char)ch) || ch == '_' || ch == '.' || ch == '-');
        if (!ok) return set_error(db, "bad collection name '" + name + "': use 1..200 characters of [A-Za-z0-9_.-]");
    }
    CollectionSpec spec;
    if (f.size() > 1 && !f[1].empty()) {
        const long d = std::strtol(f[1].c_str(), nullptr, 10);
        if (d <= 0) return set_error(db, "bad dimension '" + f[1] + "'");
        spec.dim = (uint32_t)d;
    }
    if (f.size() > 2 && !f[2].empty()) {
        const std::string m = lower(f[2]);
        if (m == "euclidean" || m == "l2") spec.metric = VROD_EUCLIDEAN;
        else if (m == "cosine" || m == "cos") spec.metric = VROD_COSINE;
        else return set_error(db, "unknown metric '" + f[2] + "' (euclidean | cosine)");
    }
    if (f.size() > 3 && !f[3].empty()) {
        const long long cap = std::strtoll(f[3].c_str(), nullptr, 10);
        if (cap <= 0) return set_error(db, "bad capacity '" + f[3] + "'");
        spec.capacity = (uint64_t)cap;
    }
    vrod_collection *c = nullptr;
    if (db->pending.count(name) || vrod_collection_get(db->ctx(), name.c_str(), &c) == VROD_OK)
        return set_error(db, "collection '" + name + "' already exists");
    db->last = CommandResult{};
    if (spec.dim) {
        if (!api(db, vrod_collection_create(db->ctx(), name.c_str(), *spec.dim, spec.metric, spec.capacity, &c))) return;
    } else {
        db->pending[name] = spec;   // the first INSERT fixes the dimension
    }
    db->payloads[name];
    db->dirty = true;
    std::printf("created %s\n", name.c_str());
}

void DropCollectionCommand::execute() const {
    if (!collection_name) return set_error(db, "DROP needs a collection name in the argument");
    db->last = CommandResult{};
    if (db->pending.erase(*collection_name) == 0 && !api(db, vrod_collection_drop(db->ctx(), collection_name->c_str()))) return;
    db->payloads.erase(*collection_name);
    db->dirty = true;
    {   // forget its files too, if this database lives in a directory
        std::error_code ec;
        if (!db->path.empty()) {
            std::filesystem::remove(db->path / (*collection_name + ".vrc"), ec);
            std::filesystem::remove(db->path / (*collection_name + ".payloads"), ec);
        }
    }
    std::printf("dropped %s\n", collection_name->c_str());
}

void ListCollectionsCommand::execute() const {
    db->last = CommandResult{};
    size_t need = 0;
    if (!api(db, vrod_collection_list(db->ctx(), nullptr, 0, &need))) return;
    std::string buf(need, '\0');
    if (!api(db, vrod_collection_list(db->ctx(), buf.data(), need, &need))) return;
    buf.resize(need ? need - 1 : 0);
    if (!buf.empty()) db->last.names = split(buf, '\n');
    for (const auto &kv : db->pending) db->last.names.push_back(kv.first);
    std::sort(db->last.names.begin(), db->last.names.end());
    for (const std::string &n : db->last.names) std::printf("%s\n", n.c_str());
}

// INSERT -c coll -a f32,f32,...[;payload]
void InsertCommand::execute() const {
    if (!collection_name || !arg) return set_error(db, "INSERT needs a collection and a record argument");
    Record rec;
    std::string err;
    if (!parse_record(*arg, &rec, &err)) return set_error(db, "INSERT: " + err);
    vrod_collection *c = resolve(db, *collection_name, (uint32_t)rec.vec.size(), 0);
    if (!c) return;
    uint32_t dim = 0;
    vrod_collection_info(c, &dim, nullptr, nullptr, nullptr);
    if (dim != rec.vec.size())
        return set_error(db, "INSERT: record has " + std::to_string(rec.vec.size()) + " components, collection has " +
                                 std::to_string(dim));
    uint64_t first = 0;
    if (!api(db, vrod_collection_insert(c, rec.vec.data(), 1, &first))) return;
    auto &pl = db->payloads[*collection_name];
    if (pl.size() <= first) pl.resize(first + 1);
    pl[first] = rec.payload;
    db->last = CommandResult{};
    db->last.first_id = first;
    db->last.inserted = 1;
    db->dirty = true;
    std::printf("inserted id %llu\n", (unsigned long long)first);
}

// BULKINSERT -c coll -a path   (file of `f32,...;payload` lines, src/utils/embeddings.rs:52-62)
void BulkInsertCommand::execute() const {
    if (!collection_name || !arg) return set_error(db, "BULKINSERT needs a collection and a file path");
    std::vector<float> rows;
    std::vector<std::string> payloads;
    uint32_t dim = 0;
    std::string err;
    if (!read_records_file(*arg, &rows, &payloads, &dim, &err)) return set_error(db, "BULKINSERT: " + err);
    vrod_collection *c = resolve(db, *collection_name, dim, payloads.size());
    if (!c) return;
    uint32_t cdim = 0;
    vrod_collection_info(c, &cdim, nullptr, nullptr, nullptr);
    if (cdim != dim)
        return set_error(db, "BULKINSERT: file has dimension " + std::to_string(dim) + ", collection has " + std::to_string(cdim));
    uint64_t first = 0;
    if (!api(db, vrod_collection_insert(c, rows.data(), payloads.size(), &first))) return;
    auto &pl = db->payloads[*collection_name];
    if (pl.size() < first + payloads.size()) pl.resize(first + payloads.size());
    std::move(payloads.begin(), payloads.end(), pl.begin() + (std::ptrdiff_t)first);
    db->last = CommandResult{};
    db->last.first_id = first;
    db->last.inserted = pl.size() - first;
    db->dirty = true;
    std::printf("inserted %llu records, first id %llu\n", (unsigned long long)db->last.inserted, (unsigned long long)first);
}

// SEARCH -c coll -a k;f32,f32,...      <- reference src/command/types.rs:114-119 (empty body)
// SEARCH -c coll -a k;@FILE            many queries at once: FILE holds one query per line in the record format
//                                      (`f32,...[;label]`, src/utils/embeddings.rs:61); they go to the library as ONE
//                                      vrod_collection_search call, i.e. through the batched tensor-core path
void SearchCommand::execute() const {
    if (!collection_name || !arg) return set_error(db, "SEARCH needs a collection and a 'k;f32,f32,...' argument");
    const size_t semi = arg->find(';');
    if (semi == std::string::npos) return set_error(db, "SEARCH argument must be 'k;f32,f32,...' or 'k;@FILE'");
    char *stop = nullptr;
    const long k = std::strtol(arg->c_str(), &stop, 10);
    if (stop != arg->c_str() + semi || k <= 0 || k > (long)VROD_MAX_K)
        return set_error(db, "SEARCH: k must be an integer in [1, " + std::to_string(VROD_MAX_K) + "]");
    std::vector<float> q;
    std::vector<std::string> labels;
    std::string err;
    uint32_t qdim = 0;
    const std::string rest = arg->substr(semi + 1);
    const bool many = !rest.empty() && rest[0] == '@';
    if (many) {
        if (!read_records_file(rest.substr(1), &q, &labels, &qdim, &err)) return set_error(db, "SEARCH: " + err);
    } else {
        if (!parse_vector(rest, &q, &err)) return set_error(db, "SEARCH: " + err);
        qdim = (uint32_t)q.size();
    }
    vrod_collection *c = resolve(db, *collection_name, std::nullopt, 0);
    if (!c) return;
    uint32_t dim = 0;
    vrod_collection_info(c, &dim, nullptr, nullptr, nullptr);
    if (dim != qdim)
        return set_error(db, "SEARCH: query has " + std::to_string(qdim) + " components, collection has " + std::to_string(dim));
    const size_t b = q.size() / dim;
    if (b > 65536) return set_error(db, "SEARCH: more than 65536 queries in one file");
    std::vector<uint64_t> ids(b * (size_t)k);
    std::vector<float> dist(b * (size_t)k);
    if (!api(db, vrod_collection_search(c, q.data(), (uint32_t)b, (uint32_t)k, ids.data(), dist.data()))) return;
    const auto &pl = db->payloads[*collection_name];
    CommandResult res;
    for (size_t qi = 0; qi < b; ++qi) {
        // many queries: a header line per query -- "# query <index> <label>" -- then its hits
        if (many) std::printf("# query %zu\t%s\n", qi, qi < labels.size() ? labels[qi].c_str() : "");
        for (size_t i = 0; i < (size_t)k; ++i) {
            const uint64_t id = ids[qi * (size_t)k + i];
            if (id == VROD_PAD_ID) break;   // fewer than k rows: the tail is padding
            const float d = dist[qi * (size_t)k + i];
            const std::string word = id < pl.size() ? pl[id] : std::string();
            // rank <TAB> id <TAB> distance (round-trip precision) <TAB> payload
            std::printf("%zu\t%llu\t%.9g\t%s\n", i + 1, (unsigned long long)id, (double)d, word.c_str());
            res.ids.push_back(id);
            res.dist.push_back(d);
            res.payload.push_back(word);
        }
    }
    db->last = std::move(res);   // all hits in query order (a single query: its k hits)
}

void TruncateWalCommand::execute() const { unsupported(db, "TRUNCATEWAL"); }
void UpdateCommand::execute() const { unsupported(db, "UPDATE"); }
void DeleteCommand::execute() const { unsupported(db, "DELETE"); }
void SearchSimilarCommand::execute() const { unsupported(db, "SEARCHSIMILAR"); }
void ReindexCommand::execute() const { unsupported(db, "REINDEX"); }
void UnrecognizedCommand::execute() const { set_error(db, "Unrecognized command: " + arg.value_or("")); }

std::unique_ptr<Command> CommandBuilder::build(OptStr collection, const std::string &command, OptStr arg) {
    std::string up = command;
    std::transform(up.begin(), up.end(), up.begin(), [](unsigned char c) { return (char)std::toupper(c); });
    std::unique_ptr<Command> cmd;
    auto with = [&](Command *c, OptStr coll, OptStr a) {
        c->db = db_;
        c->collection_name = std::move(coll);
        c->arg = std::move(a);
        cmd.reset(c);
    };
    if (up == "CREATE") with(new CreateCollectionCommand, arg, std::nullopt);
    else if (up == "DROP") with(new DropCollectionCommand, arg, std::nullopt);
    else if (up == "LISTCOLLECTIONS") with(new ListCollectionsCommand, std::nullopt, std::nullopt);
    else if (up == "TRUNCATEWAL") with(new TruncateWalCommand, collection, std::nullopt);
    else if (up == "INSERT") with(new InsertCommand, collection, arg);
    else if (up == "BULKINSERT") with(new BulkInsertCommand, collection, arg);
    else if (up == "UPDATE") with(new UpdateCommand, collection, arg);
    else if (up == "DELETE") with(new DeleteCommand, collection, arg);
    else if (up == "SEARCH") with(new SearchCommand, collection, arg);
    else if (up == "SEARCHSIMILAR") with(new SearchSimilarCommand, collection, arg);
    else if (up == "REINDEX") with(new ReindexCommand, collection, std::nullopt);
    else throw CommandBuilderError(command);
    return cmd;
}

}  // namespace vrod
