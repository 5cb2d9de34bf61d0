#!/usr/bin/env python3
"""One arena close to the int32 limit (default 450 000 x 4 KiB = 1.84 GB) through the scan, against the oracle."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tosem-2021-replication_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
import tosemscan as ts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 450000
c = ts.gen_corpus(0x7053454D0009, n, 0, 4096, n_groups=7)
print("arena %.3f GB, %d files" % (int(c.off[-1]) / 1e9, n))
sc = ts.Scanner(0, int(c.off[-1]) + 4096, n, 16)
t0 = time.time()
got = sc.scan(c, 0)
t1 = time.time()
want = orc.scan(c.arena, c.off, c.len, c.ext, c.grp, c.n_groups, events=False) if "events" in orc.scan.__code__.co_varnames else orc.scan(c.arena, c.off, c.len, c.ext, c.grp, c.n_groups)
t2 = time.time()
ok = all(np.array_equal(got["stats"][f], want["stats"][f]) for f in ("n_lines", "n_assert", "n_headers", "n_fixture", "digest"))
ok &= np.array_equal(got["group_counts"], want["group_counts"]) and np.array_equal(got["global_counts"], want["global_counts"])
print("GPU e2e %.3f s, oracle %.1f s, totals %s, identical: %s" % (t1 - t0, t2 - t1, got["totals"].tolist(), ok))
sys.exit(0 if ok else 1)
