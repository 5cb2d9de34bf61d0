"""`tosem-scan history`: S8 on a real git repository.  The package ships no repository, so
the reader of the object store (host/git_store.hpp: loose objects, packfiles with delta chains, refs) and the churn rows
are pinned against `git` itself on a repository built here:

  CPU  --dry-run rows (changed test files per commit, object names, sizes, content checksums) == `git diff-tree` +
       `git cat-file`, before and after `git gc --aggressive` (loose objects vs one pack with OFS deltas);
  GPU  added / removed per (commit, file) == `git diff --minimal --numstat` (the minimal script: the same LCS)."""
import csv
import os
import shutil
import subprocess

import numpy as np
import pytest

import tosemscan as ts

HERE = os.path.dirname(os.path.abspath(__file__))
CLI = os.path.join(HERE, "..", "tosem-2021-replication_b200", "tosemscan", "tosem-scan")
EMPTY_TREE = "4b825dc642cb6eb9a060e54bf8d69288fbee4904"
pytestmark = pytest.mark.skipif(shutil.which("git") is None, reason="needs the git command line to build the repository")


def git(repo, *args, text=True):
    env = dict(os.environ, GIT_AUTHOR_NAME="a", GIT_AUTHOR_EMAIL="a@example.org", GIT_COMMITTER_NAME="c",
               GIT_COMMITTER_EMAIL="c@example.org", GIT_AUTHOR_DATE="2021-03-01T12:00:00Z", GIT_COMMITTER_DATE="2021-03-01T12:00:00Z",
               GIT_CONFIG_NOSYSTEM="1", HOME=str(repo))
    out = subprocess.run(["git", "-C", str(repo)] + list(args), capture_output=True, env=env, check=True)
    return out.stdout.decode() if text else out.stdout


def whole_lines(b):
    """LF-terminated lines only, no CR: the cases where git's notion of a line and SPEC section 2 coincide."""
    b = b.replace(b"\r", b"")
    return b[:b.rfind(b"\n") + 1] if b"\n" in b else b"x = 1\n"


def build_repo(root):
    repo = root / "repo"
    os.makedirs(repo)
    git(repo, "init", "-q", ".")
    base = ts.gen_corpus(0x715, 14, 1, pinned=False)
    files = {}
    names = ["tests/test_a.py", "tests/unit/test_b.py", "tests/unit/deep/c_test.cc", "src/core_test.cpp", "pkg/TestThing.java",
             "tests/data_test.h", "tests/helper_test.c", "tests/test_big.py", "src/main.c", "docs/readme.md", "tests/notes.txt",
             "tests/test_gone.py", "lib/testing/util.py", "tests/test_same.py"]
    for i, nm in enumerate(names):
        files[nm] = whole_lines(base.file_bytes(i)[:60000])
    files["tests/test_big.py"] = whole_lines(b"".join(base.file_bytes(i) for i in range(14))[:400000])

    def commit(msg):
        for nm, data in files.items():
            p = repo / nm
            os.makedirs(p.parent, exist_ok=True)
            p.write_bytes(data)
        git(repo, "add", "-A")
        git(repo, "commit", "-q", "-m", msg)

    (repo / "tests").mkdir()
    (repo / "tests" / "blob_test.py").write_bytes(b"\x00\x01binary\x00" * 50)           # binary: no line counts, like numstat
    commit("c1")
    for k, nm in enumerate(["tests/test_a.py", "tests/unit/deep/c_test.cc", "pkg/TestThing.java", "tests/test_big.py", "src/main.c", "tests/notes.txt"]):
        files[nm] = whole_lines(ts.gen_edit(100 + k, files[nm], 5.0))
    commit("c2")
    git(repo, "tag", "-a", "v1", "-m", "first")
    del files["tests/test_gone.py"]
    os.remove(repo / "tests" / "test_gone.py")
    files["tests/new_test.cc"] = b"TEST(New, Case) {\n  EXPECT_EQ(1, 1);\n  ASSERT_TRUE(x);\n}\n"
    files["tests/unit/test_b.py"] = whole_lines(ts.gen_edit(7, files["tests/unit/test_b.py"], 40.0))
    (repo / "tests" / "blob_test.py").write_bytes(b"\x00\x02binary\x00" * 60)
    commit("c3")
    files["tests/test_big.py"] = whole_lines(ts.gen_edit(9, files["tests/test_big.py"], 25.0))
    files["tests/helper_test.c"] = b""                                                   # emptied
    files["tests/new_test.cc"] += b"TEST(New, Other) {\n  EXPECT_NE(a, b);\n}\n"
    commit("c4")
    git(repo, "tag", "v2")
    files["tests/test_big.py"] = whole_lines(ts.gen_edit(11, files["tests/test_big.py"], 3.0)) + b"assert tail == 5\n"
    commit("c5")
    return repo


def is_test_file(path):
    return "test" in path.lower() and path.rsplit(".", 1)[-1] in ("py", "cc", "cpp", "java", "c", "h") and "." in path.rsplit("/", 1)[-1]


def fnv(b):
    h = 0xcbf29ce484222325
    for x in np.frombuffer(b, np.uint8).tolist():
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def expected_changes(repo):
    """[(commit, parent, path, old_oid, new_oid)] of the test files along the first-parent chain, from git itself."""
    chain = git(repo, "rev-list", "--first-parent", "--reverse", "HEAD").split()
    out = []
    for i, c in enumerate(chain):
        parent = chain[i - 1] if i else ""
        raw = git(repo, "diff-tree", "-r", "--no-renames", "--raw", "--no-commit-id", "--root", "-z", c) if not i else \
            git(repo, "diff-tree", "-r", "--no-renames", "--raw", "-z", parent, c)
        parts = raw.split("\0")
        for j in range(0, len(parts) - 1, 2):
            meta, path = parts[j], parts[j + 1]
            f = meta.lstrip(":").split()
            old_mode, new_mode, old, new = f[0], f[1], f[2], f[3]
            if not is_test_file(path):
                continue
            out.append((c, parent, path, "" if old_mode == "000000" else old, "" if new_mode == "000000" else new))
    return chain, out


def run_history(repo, out, *extra):
    r = subprocess.run([CLI, "history", str(repo), "--out", str(out)] + list(extra), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return list(csv.DictReader(open(out, newline=""))), r


@pytest.fixture(scope="module")
def repo(tmp_path_factory):
    return build_repo(tmp_path_factory.mktemp("hist"))


def test_object_store_reader_matches_git_loose_and_packed(repo, tmp_path):
    chain, want = expected_changes(repo)
    assert len(chain) == 5 and len(want) > 15

    def check(tag):
        rows, r = run_history(repo, tmp_path / ("dry_%s.csv" % tag), "--dry-run")
        got = sorted((x["commit"], x["parent"], x["fileName"], x["old_blob"], x["new_blob"]) for x in rows)
        assert got == sorted(want)
        for x in rows:                                       # every blob byte for byte (zlib, delta chains)
            for side in ("old", "new"):
                oid = x[side + "_blob"]
                data = git(repo, "cat-file", "blob", oid, text=False) if oid else b""
                assert int(x[side + "_size"]) == len(data) and x[side + "_fnv"] == fnv(data), (x["fileName"], side)
        assert "5 commits" in r.stderr
        return rows

    loose = check("loose")
    git(repo, "gc", "-q", "--aggressive", "--prune=now")
    packs = [f for f in os.listdir(repo / ".git" / "objects" / "pack") if f.endswith(".pack")]
    assert len(packs) == 1
    verify = git(repo, "verify-pack", "-v", os.path.join(".git", "objects", "pack", packs[0]))
    assert any(len(line.split()) == 7 for line in verify.splitlines()), "the pack holds no delta: the delta path is not exercised"
    packed = check("packed")
    assert loose == packed
    # tags (annotated: peeled through the tag object; lightweight) and a raw object name resolve to the same chain prefixes
    for rev, n in (("v1", 2), ("v2", 4), (chain[2], 3), ("refs/tags/v2", 4)):
        rows, r = run_history(repo, tmp_path / "dry_rev.csv", "--dry-run", "--rev", rev)
        assert "%d commits" % n in r.stderr, (rev, r.stderr)
    rows, r = run_history(repo, tmp_path / "dry_max.csv", "--dry-run", "--max-commits", "2")
    assert {x["commit"] for x in rows} <= set(chain[-2:])


@pytest.mark.gpu
def test_history_churn_matches_git_numstat(repo, tmp_path):
    chain, want = expected_changes(repo)
    rows, r = run_history(repo, tmp_path / "churn.csv")
    got = {(x["commit"], x["fileName"]): (int(x["added"]), int(x["removed"])) for x in rows}
    n = 0
    for i, c in enumerate(chain):
        parent = chain[i - 1] if i else EMPTY_TREE
        for line in git(repo, "diff", "--minimal", "--numstat", "--no-renames", parent, c).splitlines():
            a, d, path = line.split("\t")
            if not is_test_file(path):
                continue
            if a == "-":                                     # binary: skipped by both
                assert (c, path) not in got
                continue
            assert got[(c, path)] == (int(a), int(d)), (c, path)
            n += 1
    assert n == len(got) and n > 15
    assert "1 binary skipped" in r.stderr or "2 binary skipped" in r.stderr
    # the assertion lines among the changed lines of one known commit (c4 appends a TEST with one EXPECT_NE)
    c4 = [x for x in rows if x["commit"] == chain[3] and x["fileName"] == "tests/new_test.cc"][0]
    assert (c4["added"], c4["removed"], c4["added_assert"], c4["hunks_add"]) == ("3", "0", "1", "1")
    # per-commit totals on stdout
    tot = {l.split(",")[0]: l.strip().split(",")[1:] for l in r.stdout.splitlines()[1:]}
    for c in chain:
        mine = [v for (cc, _), v in got.items() if cc == c]
        assert [int(v) for v in tot[c]] == [len(mine), sum(a + d for a, d in mine), sum(a for a, _ in mine), sum(d for _, d in mine)]


@pytest.mark.gpu
def test_release_matrix_from_tags_equals_the_matrix_of_checkouts(repo, tmp_path):
    """`releases --git` reads the trees of the tags from the object store; the matrix must equal the one `releases`
    builds from work trees of the same revisions (SPEC section 11 is unchanged: only where the bytes come from differs)."""
    roots = []
    for tag in ("v1", "v2", "HEAD"):
        d = tmp_path / ("co_" + tag)
        os.makedirs(d)
        tar = git(repo, "archive", "--format=tar", tag, text=False)
        subprocess.run(["tar", "-x", "-C", str(d)], input=tar, check=True)
        roots.append("%s=%s" % (d, tag))
    a, b = tmp_path / "from_dirs.csv", tmp_path / "from_git.csv"
    r1 = subprocess.run([CLI, "releases"] + roots + ["--out", str(a)], capture_output=True, text=True)
    r2 = subprocess.run([CLI, "releases", "--git", str(repo), "v1", "v2", "HEAD", "--out", str(b)], capture_output=True, text=True)
    assert r1.returncode == 0 and r2.returncode == 0, (r1.stderr, r2.stderr)
    assert open(a, "rb").read() == open(b, "rb").read() and r1.stdout == r2.stdout
    rows = list(csv.reader(open(b, newline="")))
    assert rows[0][:2] == ["Id", "FileName"] and rows[0][2:5] == ["v1", "v2", "HEAD"]
    by = {x[1]: x for x in rows[1:]}
    assert by["tests/test_gone.py"][2:5] == ["tests/test_gone.py", "", ""]              # deleted after v1
    assert by["tests/new_test.cc"][2:5] == ["", "tests/new_test.cc", "tests/new_test.cc"] and by["tests/new_test.cc"][5] == "3"
    # no revisions named: every tag, oldest first
    r3 = subprocess.run([CLI, "releases", "--git", str(repo), "--out", str(tmp_path / "tags.csv")], capture_output=True, text=True)
    assert r3.returncode == 0, r3.stderr
    assert next(csv.reader(open(tmp_path / "tags.csv", newline="")))[2:4] == ["v1", "v2"]
