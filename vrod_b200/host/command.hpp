// command.hpp -- host-side mirror of the reference's command layer, C++17 (the reference is Rust;
// there is no Rust toolchain in this image).
//
//   trait Command { fn execute(&self); }                      src/command/types.rs:5-7
//   twelve command structs, each { db, collection_name?, arg? } src/command/types.rs:9-154
//   CommandBuilder::build(collection, command, arg)           src/command/builder.rs:22-81
//
// `std::shared_ptr<Database>` plays Rc<RefCell<Database>> (types.rs:10): single-threaded sharing,
// no re-entrancy.  execute() returns nothing, exactly like the trait; the outcome is left in
// db->last (database.hpp) and printed to stdout by the commands that produce output.
#pragma once
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>

#include "database.hpp"

namespace vrod {

using Db = std::shared_ptr<Database>;
using OptStr = std::optional<std::string>;

struct Command {
    virtual ~Command() = default;
    virtual void execute() const = 0;
    virtual const char *type_name() const = 0;          // for tests / --describe
    virtual std::string describe() const;               // "Type{collection_name=..,arg=..}"
    Db db;
    OptStr collection_name;   // `target` for TruncateWal
    OptStr arg;
};

#define VROD_DECLARE_COMMAND(Name)                         \
    struct Name : Command {                                \
        void execute() const override;                     \
        const char *type_name() const override { return #Name; } \
    };
VROD_DECLARE_COMMAND(CreateCollectionCommand)   // types.rs:9-19    arg = name[;dim[;metric[;capacity]]]
VROD_DECLARE_COMMAND(DropCollectionCommand)     // types.rs:21-32   arg = name
VROD_DECLARE_COMMAND(ListCollectionsCommand)    // types.rs:34-43
VROD_DECLARE_COMMAND(TruncateWalCommand)        // types.rs:44-54   out of scope: reports "not supported"
VROD_DECLARE_COMMAND(InsertCommand)             // types.rs:56-67   arg = f32,f32,...[;payload]
VROD_DECLARE_COMMAND(BulkInsertCommand)         // types.rs:69-80   arg = path of a records file
VROD_DECLARE_COMMAND(UpdateCommand)             // types.rs:82-93   out of scope
VROD_DECLARE_COMMAND(DeleteCommand)             // types.rs:95-106  out of scope
VROD_DECLARE_COMMAND(SearchCommand)             // types.rs:108-119 arg = k;f32,f32,... | k;@FILE   <- the hot path
VROD_DECLARE_COMMAND(SearchSimilarCommand)      // types.rs:121-132 out of scope (unspecified upstream)
VROD_DECLARE_COMMAND(ReindexCommand)            // types.rs:134-144 out of scope (exact scan has no index)
VROD_DECLARE_COMMAND(UnrecognizedCommand)       // types.rs:146-154
#undef VROD_DECLARE_COMMAND

// builder.rs:10-15: #[error("Unrecognized command: {0}")]
struct CommandBuilderError : std::runtime_error {
    explicit CommandBuilderError(const std::string &command)
        : std::runtime_error("Unrecognized command: " + command), command(command) {}
    std::string command;
};

class CommandBuilder {
  public:
    explicit CommandBuilder(Db db) : db_(std::move(db)) {}
    // Case-insensitive command name (builder.rs:29 to_uppercase).  CREATE / DROP take the collection
    // name from `arg`; TRUNCATEWAL takes `collection` as its target; LISTCOLLECTIONS takes nothing;
    // INSERT, BULKINSERT, UPDATE, DELETE, SEARCH, SEARCHSIMILAR take collection + arg; REINDEX takes
    // collection.  Anything else throws CommandBuilderError.
    std::unique_ptr<Command> build(OptStr collection, const std::string &command, OptStr arg);

  private:
    Db db_;
};

}  // namespace vrod
