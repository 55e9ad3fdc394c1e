"""GPU test of the caller side of the boundary: CREATE / BULKINSERT / INSERT / SEARCH / LISTCOLLECTIONS /
DROP through the `vrod` CLI (C++ mirror of the reference's command layer), answers against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import assert_same

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "vrod_b200", "host", "vrod")


def fmt(v):
    return ",".join(repr(float(np.float32(x))) if np.float32(float(repr(float(np.float32(x))))) == np.float32(x) else "%.9g" % x
                    for x in v)


def parse_hits(lines):
    ids, dist, words = [], [], []
    for ln in lines:
        rank, i, d, w = ln.split("\t")
        ids.append(int(i))
        dist.append(np.float32(float(d)))
        words.append(w)
    return np.array(ids, dtype=np.uint64), np.array(dist, dtype=np.float32), words


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_cli_end_to_end(tmp_path, oracle, metric):
    n, d, k = 500, 24, 7
    X = oracle.fill(n + 1, d, 99)
    words = [f"w{i}" for i in range(n + 1)]
    path = tmp_path / "alice_embeddings.txt"
    with open(path, "w") as f:                      # the reference's record format, embeddings.rs:61
        for i in range(n):
            f.write("%s;%s\n" % (",".join("%.9g" % x for x in X[i]), words[i]))
    q = oracle.fill(1, d, 100)[0]
    script = "\n".join([
        f"CREATE - words;;{metric}",                # dimension fixed by the first insert
        f"BULKINSERT words {path}",
        "INSERT words %s;%s" % (",".join("%.9g" % x for x in X[n]), words[n]),
        "SEARCH words %d;%s" % (k, ",".join("%.9g" % x for x in q)),
        "LISTCOLLECTIONS",
        "SEARCH words 2000;%s" % ",".join("%.9g" % x for x in q),
        "SEARCH words 3;1,2,3",
        "SEARCH nope 3;1,2,3",
        "DROP - words",
        "LISTCOLLECTIONS",
    ]) + "\n"
    r = subprocess.run([CLI, "--script", "-"], input=script, capture_output=True, text=True, timeout=300)
    out = r.stdout.splitlines()
    assert out[0] == "created words" and out[1] == f"inserted {n} records, first id 0" and out[2] == f"inserted id {n}", r.stderr
    ids, dist, ws = parse_hits(out[3:3 + k])
    rid, rdist = oracle.search(X, q, k, 0 if metric == "euclidean" else 1)
    assert_same(ids, dist, rid[0], rdist[0], "CLI SEARCH")
    assert ws == [words[int(i)] for i in rid[0]]
    assert out[3 + k] == "words"
    assert out[4 + k:] == ["dropped words"]
    assert r.returncode == 1                          # some commands failed, as intended:
    assert "k must be an integer in [1, 1024]" in r.stderr
    assert "query has 3 components, collection has 24" in r.stderr
    assert "no collection 'nope'" in r.stderr


def test_database_directory_persists_across_processes(tmp_path, oracle):
    """--init-database, then one command per process with the reference's flags (-d -c -e -a): the collection
    created and filled by earlier processes is loaded from DIR and searched by a later one."""
    n, d, k = 300, 16, 5
    X = oracle.fill(n, d, 7)
    recs = tmp_path / "recs.txt"
    with open(recs, "w") as f:
        for i in range(n):
            f.write("%s;word%d\n" % (",".join("%.9g" % x for x in X[i]), i))
    db = tmp_path / "mydb"

    def cli(*args):
        return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=300)

    assert cli("-i", str(tmp_path), "-n", "mydb").returncode == 0
    r = cli("-d", str(db), "-e", "CREATE", "-a", "words;;cosine")
    assert r.returncode == 0 and "created words" in r.stdout, r.stderr
    r = cli("-d", str(db), "-c", "words", "-e", "BULKINSERT", "-a", str(recs))
    assert r.returncode == 0 and f"inserted {n} records" in r.stdout, r.stderr
    assert sorted(os.listdir(db)) == ["vr_config", "vr_wal", "words.payloads", "words.vrc"]
    q = oracle.fill(1, d, 8)[0]
    r = cli("--database", str(db), "--collection", "words", "--execute", "search", "--command-arg",
            "%d;%s" % (k, ",".join("%.9g" % x for x in q)))
    assert r.returncode == 0, r.stderr
    ids, dist, words = parse_hits(r.stdout.splitlines())
    rid, rdist = oracle.search(X, q, k, 1)
    assert_same(ids, dist, rid[0], rdist[0], "SEARCH after reload")
    assert words == [f"word{int(i)}" for i in rid[0]]
    r = cli("-d", str(db), "-e", "LISTCOLLECTIONS")
    assert r.stdout.split() == ["words"]
    assert cli("-d", str(db), "-e", "DROP", "-a", "words").returncode == 0
    assert sorted(os.listdir(db)) == ["vr_config", "vr_wal"]
    assert cli("-d", str(db), "-e", "LISTCOLLECTIONS").stdout.strip() == ""


def test_cli_search_with_a_file_of_queries(tmp_path, oracle):
    """SEARCH coll 'k;@FILE': one query per line of FILE in the record format; the host layer hands them to the
    library as ONE call (200 queries over 30k rows: the batched tensor-core path) and prints the hits per query."""
    n, d, k, b = 30000, 32, 5, 200
    X = oracle.fill(n, d, 61)
    Q = oracle.fill(b, d, 62)
    rows, queries = tmp_path / "rows.txt", tmp_path / "queries.txt"
    with open(rows, "w") as f:
        for i in range(n):
            f.write("%s;w%d\n" % (",".join("%.9g" % x for x in X[i]), i))
    with open(queries, "w") as f:
        for i in range(b):
            f.write("%s;q%d\n" % (",".join("%.9g" % x for x in Q[i]), i))
    script = f"CREATE - words;;euclidean\nBULKINSERT words {rows}\nSEARCH words {k};@{queries}\n"
    r = subprocess.run([CLI, "--script", "-"], input=script, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout.splitlines()[2:]
    assert len(out) == b * (k + 1)
    rid, rdist = oracle.search(X, Q, k, 0)
    for qi in range(b):
        block = out[qi * (k + 1):(qi + 1) * (k + 1)]
        assert block[0] == f"# query {qi}\tq{qi}"
        ids, dist, ws = parse_hits(block[1:])
        assert_same(ids, dist, rid[qi], rdist[qi], f"query {qi}")
        assert ws == [f"w{int(i)}" for i in rid[qi]]
