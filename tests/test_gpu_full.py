"""GPU parity at the sizes and on the bytes BASELINE.json names:

  C1  the study's own test files (tests/golden/c1_testfiles.npz, made by tools/make_golden.py from
      /root/reference/src/**) and the hazard files of SURVEY.md section 8d, whole arena, events included;
  C2  100 000 x 4 KiB, every file against the oracle;
  C4  100 000 Zipf-sized files (1.14 GB), every file against the oracle;
  C5  50 000 revision pairs: added / removed / hunks / changed assertion lines of every pair against the
      oracle's serial Myers script, 5 000 pairs against the oracle's O(n*m) LCS table.

The oracle runs on all host cores here (oracle/orc_mt.c); it is the checker, never the thing measured."""
import json
import os

import numpy as np
import pytest

import corpus_util as cu
import orc
import tosemscan as ts

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLAGS = ts.SCAN_ASSERT_EVENTS | ts.SCAN_HEADER_EVENTS


def compare_scan(got, want, events):
    for f in ("n_lines", "n_assert", "n_headers", "n_fixture", "digest"):
        bad = np.nonzero(got["stats"][f] != want["stats"][f])[0]
        assert bad.size == 0, (f, bad[:10], got["stats"][bad[:5]], want["stats"][bad[:5]])
    assert np.array_equal(got["group_counts"], want["group_counts"])
    assert np.array_equal(got["global_counts"], want["global_counts"])
    if events:
        for k in ("assert_events", "header_events"):
            a, b = got[k], want[k]
            assert len(a) == len(b), (k, len(a), len(b))
            for f in a.dtype.names:
                bad = np.nonzero(a[f] != b[f])[0]
                assert bad.size == 0, (k, f, a[bad[:5]], b[bad[:5]])


@pytest.mark.parametrize("name", ["c1_testfiles.npz", "c1_hazard_files.npz"])
def test_c1_real_bytes_whole_arena(name):
    files, exts, grps, n_groups = cu.load_fixture(os.path.join(GOLD, name))
    c = ts.pack(files, exts, grps, n_groups)
    sc = ts.Scanner(0, int(c.off[-1]) + 4096, c.n_files, 16)
    got = sc.scan(c, FLAGS)
    want = orc.scan(c.arena, c.off, c.len, c.ext, c.grp, n_groups)
    compare_scan(got, want, True)
    got2 = sc.scan(c, 0)                                   # and the path the bench times (no events)
    assert np.array_equal(got2["stats"], want["stats"]) and np.array_equal(got2["group_counts"], want["group_counts"])
    if name == "c1_testfiles.npz":                          # the committed summary of config C1, from the GPU's numbers
        s = json.load(open(os.path.join(GOLD, "c1_summary.json")))
        st = got["stats"]
        assert c.n_files == s["n_files"] and c.source_bytes == s["bytes"]
        assert got["totals"].tolist() == [s["n_lines"], s["n_assert"], s["n_headers"], s["n_fixture"]]
        assert "%016x" % int(np.bitwise_xor.reduce(st["digest"])) == s["digest_xor"]
        assert {ts.category_name(i) or "''": int(v) for i, v in enumerate(got["global_counts"]) if v} == s["global_counts"]
        assert [int(got["group_counts"][g].sum()) for g in range(n_groups)] == [s["per_project_assert"][p] for p in s["projects"]]
    sc.close()


def _oracle_all_cores(c):
    mt = orc.MtScanner(0, max_groups=max(c.n_groups, 1))
    try:
        want = mt.scan(c.arena, c.off, c.len, c.ext, c.grp, c.n_groups)
        return {k: v.copy() for k, v in want.items()}
    finally:
        mt.close()


def test_full_c2_every_file_against_the_oracle():
    c = ts.gen_corpus(0x7053454D0002, 100000, 0, 4096, n_groups=9)
    sc = ts.Scanner(0, int(c.off[-1]) + 4096, c.n_files, 16)
    got = sc.scan(c, 0)
    compare_scan(got, _oracle_all_cores(c), False)
    sc.upload(c)                                            # the resident path (what `value` times)
    sc.scan_resident(0)
    res = sc.download(0)
    assert np.array_equal(res["stats"], got["stats"]) and np.array_equal(res["group_counts"], got["group_counts"])
    sc.close()


def test_full_c4_zipf_every_file_against_the_oracle():
    c = ts.gen_corpus(0x7053454D0004, 100000, 1, n_groups=9)
    assert c.source_bytes > 1_000_000_000 and int(c.len.max()) > 500_000
    sc = ts.Scanner(0, int(c.off[-1]) + 4096, c.n_files, 16)
    got = sc.scan(c, 0)
    compare_scan(got, _oracle_all_cores(c), False)
    sc.close()


def test_full_c5_pairs_against_the_oracle():
    n = 50000
    a, b = ts.gen_pairs(0x7053454D0005, n)
    assert a.n_files == n and 150e6 < a.source_bytes < 400e6
    sc = ts.Scanner(0, 1 << 20, 16, 1)
    add, rem, det = sc.diff_pairs(a, b, detail=True)
    wadd, wrem, wdet = orc.diff_pairs_detail((a.arena, a.off, a.len, a.ext), (b.arena, b.off, b.len, b.ext))
    bad = np.nonzero((add != wadd) | (rem != wrem))[0]
    assert bad.size == 0, (bad[:5], add[bad[:5]], wadd[bad[:5]])
    for f in det.dtype.names:
        bad = np.nonzero(det[f] != wdet[f])[0]
        assert bad.size == 0, (f, bad[:5], det[bad[:5]], wdet[bad[:5]])
    # cloc = added + removed (ML-Testing-v1.xlsx!projects:R1); added - removed = change in line count
    nl_a = orc.scan(a.arena, a.off, a.len, a.ext, a.grp, 1, events=False)["stats"]["n_lines"].astype(np.int64)
    nl_b = orc.scan(b.arena, b.off, b.len, b.ext, b.grp, 1, events=False)["stats"]["n_lines"].astype(np.int64)
    assert np.array_equal(add - rem, nl_b - nl_a)
    # 5 000 pairs against the LCS table (independent of the Myers script)
    idx = np.arange(0, n, 10)
    sa = ts.pack([a.file_bytes(int(i)) for i in idx], a.ext[idx])
    sb = ts.pack([b.file_bytes(int(i)) for i in idx], b.ext[idx])
    dadd, drem = orc.diff_pairs((sa.arena, sa.off, sa.len), (sb.arena, sb.off, sb.len))
    assert np.array_equal(add[idx], dadd) and np.array_equal(rem[idx], drem)
    sc.close()
