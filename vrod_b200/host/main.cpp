// main.cpp -- `vrod` CLI with the reference's flag surface (src/main.rs:10-34), driving the command
// layer the reference never wired to main (src/main.rs:64-74 is commented out).
//
//   -i/--init-database PATH  -n/--init-database-name NAME   create PATH/NAME with vr_config + vr_wal
//   -d/--database DIR  -c/--collection NAME  -e/--execute COMMAND  -a/--command-arg ARG
//   -g/--generate-embeddings AMOUNT   dev-only in the reference (fastembed); not provided here
// With -d DIR the collections are loaded from DIR at start and written back after the commands (rows as
// <name>.vrc, payloads as <name>.payloads, the list in vr_config).  Additions:
//   --script FILE|-   one command per line: COMMAND <collection|-> [ARG...rest of line]
//   --describe        build the command and print its type and fields instead of executing it
//   --device N        CUDA ordinal (default 0)
//   --devices A,B,..  several GPUs driven by this one process: every collection is row-sharded over them
//                     (vrod_ctx_create_multi), a SEARCH scans all shards at once
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "command.hpp"

using namespace vrod;

static int usage() {
    std::fputs("Usage: vrod [-i PATH -n NAME] [-d DIR] [-c COLLECTION_NAME] [-e COMMAND] [-a COMMAND_ARG]\n"
               "            [--script FILE|-] [--describe] [--device N | --devices A,B,...]\n", stderr);
    return 2;
}

int main(int argc, char **argv) {
    if (argc == 1) return usage();   // #[command(arg_required_else_help(true))], main.rs:11
    OptStr init_db, init_name, database, collection, execute, command_arg, script;
    bool describe = false;
    std::vector<int> devices{0};
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](OptStr &dst) {
            if (i + 1 >= argc) { std::fprintf(stderr, "error: %s needs a value\n", a.c_str()); std::exit(2); }
            dst = argv[++i];
        };
        if (a == "-i" || a == "--init-database") val(init_db);
        else if (a == "-n" || a == "--init-database-name") val(init_name);
        else if (a == "-d" || a == "--database") val(database);
        else if (a == "-c" || a == "--collection") val(collection);
        else if (a == "-e" || a == "--execute") val(execute);
        else if (a == "-a" || a == "--command-arg") val(command_arg);
        else if (a == "--script") val(script);
        else if (a == "--describe") describe = true;
        else if (a == "--device") { OptStr d; val(d); devices = {std::atoi(d->c_str())}; }
        else if (a == "--devices") {
            OptStr d;
            val(d);
            devices.clear();
            std::stringstream ss(*d);
            std::string tok;
            while (std::getline(ss, tok, ',')) {
                if (tok.empty() || tok.find_first_not_of("0123456789") != std::string::npos) {
                    std::fprintf(stderr, "error: --devices takes a comma-separated list of CUDA ordinals, got '%s'\n", d->c_str());
                    return 2;
                }
                devices.push_back(std::atoi(tok.c_str()));
            }
            if (devices.empty()) { std::fputs("error: --devices needs at least one ordinal\n", stderr); return 2; }
        }
        else if (a == "-g" || a == "--generate-embeddings") {
            std::fputs("error: --generate-embeddings is the reference's dev-only fastembed path; not part of this build\n", stderr);
            return 2;
        } else if (a == "-h" || a == "--help") { usage(); return 0; }
        else { std::fprintf(stderr, "error: unexpected argument '%s'\n", a.c_str()); return usage(); }
    }
    try {
        if (init_db) {   // main.rs:51-62
            if (!init_name) {
                std::fputs("Error: Missing '--init_database_name' flag with argument for '--init_database' flag.\n", stderr);
                return 1;
            }
            Database::create(*init_db, *init_name);
            return 0;
        }
        // -d DIR: main.rs:64-74 (commented out upstream) loads the database from DIR; without -d the database
        // lives in GPU memory for this process only
        Db db = database ? std::make_shared<Database>(Database::load(*database, devices)) : std::make_shared<Database>(devices);
        CommandBuilder builder(db);
        int rc = 0;
        auto run = [&](OptStr coll, const std::string &cmd, OptStr arg) {
            std::unique_ptr<Command> c;
            try {
                c = builder.build(std::move(coll), cmd, std::move(arg));
            } catch (const CommandBuilderError &e) {
                std::fprintf(stderr, "Error: %s\n", e.what());
                rc = 1;
                return;
            }
            if (describe) { std::printf("%s\n", c->describe().c_str()); return; }
            c->execute();
            if (!db->last.ok) rc = 1;
        };
        if (execute) run(collection, *execute, command_arg);
        if (script) {
            std::ifstream file;
            std::istream *in = &std::cin;
            if (*script != "-") {
                file.open(*script);
                if (!file) { std::fprintf(stderr, "error: cannot open script '%s'\n", script->c_str()); return 1; }
                in = &file;
            }
            std::string line;
            while (std::getline(*in, line)) {
                if (line.empty() || line[0] == '#') continue;
                std::istringstream ss(line);
                std::string cmd, coll, rest;
                ss >> cmd >> coll;
                std::getline(ss, rest);
                const size_t p = rest.find_first_not_of(" \t");
                rest = p == std::string::npos ? "" : rest.substr(p);
                run(coll.empty() || coll == "-" ? OptStr() : OptStr(coll), cmd, rest.empty() ? OptStr() : OptStr(rest));
            }
        }
        if (database && db->dirty && !describe) db->save();
        return rc;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
}
