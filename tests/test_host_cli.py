"""CPU tests of the C++ host layer (vrod_b200/host): the command dispatch mirrors the reference's
CommandBuilder (src/command/builder.rs:22-81) and the --init-database flow mirrors
src/main.rs:51-62 + src/database/setup.rs:3-26.  No GPU needed: --describe builds without executing."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "vrod_b200", "host", "vrod")


def run(*args, stdin=None):
    return subprocess.run([CLI, *args], capture_output=True, text=True, input=stdin, timeout=120)


# builder.rs:30-76 -- which of (collection, arg) each command keeps
ROUTING = [
    ("CREATE", "CreateCollectionCommand", "arg", None),
    ("DROP", "DropCollectionCommand", "arg", None),
    ("LISTCOLLECTIONS", "ListCollectionsCommand", None, None),
    ("TRUNCATEWAL", "TruncateWalCommand", "coll", None),
    ("INSERT", "InsertCommand", "coll", "arg"),
    ("BULKINSERT", "BulkInsertCommand", "coll", "arg"),
    ("UPDATE", "UpdateCommand", "coll", "arg"),
    ("DELETE", "DeleteCommand", "coll", "arg"),
    ("SEARCH", "SearchCommand", "coll", "arg"),
    ("SEARCHSIMILAR", "SearchSimilarCommand", "coll", "arg"),
    ("REINDEX", "ReindexCommand", "coll", None),
]


@pytest.mark.parametrize("name,typ,first,second", ROUTING)
def test_builder_routing(name, typ, first, second):
    val = {"coll": 'Some("C")', "arg": 'Some("A")', None: "None"}
    for spelling in (name, name.lower(), name.capitalize()):      # builder.rs:29 to_uppercase
        r = run("--describe", "-c", "C", "-e", spelling, "-a", "A")
        assert r.returncode == 0, r.stderr
        assert r.stdout.strip() == f"{typ}{{collection_name={val[first]},arg={val[second]}}}"


def test_unrecognized_command():
    r = run("--describe", "-e", "FROBNICATE")
    assert r.returncode == 1 and "Unrecognized command: FROBNICATE" in r.stderr   # builder.rs:12


def test_no_arguments_prints_help():
    r = run()
    assert r.returncode == 2 and "Usage" in r.stderr                              # main.rs:11


def test_init_database(tmp_path):
    r = run("-i", str(tmp_path), "-n", "db1")
    assert r.returncode == 0, r.stderr
    assert sorted(os.listdir(tmp_path / "db1")) == ["vr_config", "vr_wal"]        # setup.rs:19-23
    assert os.path.getsize(tmp_path / "db1" / "vr_wal") == 0
    r = run("-i", str(tmp_path), "-n", "db1")
    assert r.returncode == 1 and "already exists" in r.stderr                     # setup.rs:6-15
    r = run("--init-database", str(tmp_path))
    assert r.returncode == 1 and "Missing '--init_database_name' flag" in r.stderr  # main.rs:38
    r = run("-i", str(tmp_path / "missing" / "parent"), "-n", "x")                # fs::create_dir, not create_dir_all
    assert r.returncode == 1


def test_script_describe_mode():
    script = "CREATE - words;4;cosine\nsearch words 3;1,2,3,4\n# comment\nListCollections\n"
    r = run("--describe", "--script", "-", stdin=script)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == ['CreateCollectionCommand{collection_name=Some("words;4;cosine"),arg=None}',
                                     'SearchCommand{collection_name=Some("words"),arg=Some("3;1,2,3,4")}',
                                     "ListCollectionsCommand{collection_name=None,arg=None}"]


def test_devices_flag_is_validated_before_anything_runs():
    """--devices A,B,.. selects the single-process multi-GPU context (vrod_ctx_create_multi); a malformed list is a usage
    error, and --describe never touches a GPU whatever the list says."""
    r = run("--describe", "--devices", "0,1,2,3", "-c", "C", "-e", "SEARCH", "-a", "3;1,2")
    assert r.returncode == 0 and r.stdout.strip() == 'SearchCommand{collection_name=Some("C"),arg=Some("3;1,2")}'
    for bad in ("0,x", "1;2", "a"):
        r = run("--describe", "--devices", bad, "-e", "LISTCOLLECTIONS")
        assert r.returncode == 2 and "--devices takes a comma-separated list" in r.stderr, (bad, r.stderr)


def test_executing_without_a_gpu_fails_loudly():
    """There is no CPU path: a command that needs the library's context reports the CUDA error and exits non-zero (on a
    GPU box the same command succeeds -- tests/test_gpu_host_cli.py)."""
    import shutil
    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0:
        pytest.skip("a GPU is visible")
    r = run("-e", "LISTCOLLECTIONS")
    assert r.returncode != 0 and r.stderr.strip(), (r.returncode, r.stdout, r.stderr)
