#!/usr/bin/env python3
"""Small scan (Rev A and Rev B) + line records + diff + statements + reduce under compute-sanitizer (run: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tosem-2021-replication_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import corpus_util as cu
import tosemscan as ts

s = ts.Scanner(0, 1 << 24, 4096, 16)
files, exts, grps = cu.edge_corpus()
r = s.scan(ts.pack(files, exts, grps, 3), 3)
files, exts, grps = cu.fuzz_corpus(5, 120, 30000, long_lines=True)
r2 = s.scan(ts.pack(files, exts, grps, 5), 3)
c = ts.gen_corpus(3, 300, 1, n_groups=4, pinned=False)
r3 = s.scan(c, 0)
a = ts.pack([b"a\nb\nc\n", b"x\n" * 50, b""], [1, 1, 1])
b = ts.pack([b"a\nc\nd\n", b"y\n" * 40, b"q\n"], [1, 1, 1])
print(s.diff_pairs(a, b, detail=True))
far_o = [b"".join(b"o%d\n" % i for i in range(60)), b"head\n" * 10 + b"".join(b"m%d\n" % i for i in range(2500)) + b"tail\n"]
far_n = [b"".join(b"n%d\n" % i for i in range(50)), b"head\n" * 10 + b"x\n" + b"".join(b"m%d\n" % i for i in range(2500)) + b"y\ntail\n"]
print(s.diff_pairs(ts.pack(far_o, [1, 1]), ts.pack(far_n, [1, 1]), detail=True))   # the left-over path: D > 31, middle > 4 096 lines
print([x[:4] for x in s.line_hashes(ts.pack(files[:40], exts[:40]), ngram=3)])
r4 = s.scan(ts.pack(files, exts, grps, 5), 3 | ts.SCAN_REV_B)
print(s.statements(c)[0][-1])
fl = (np.random.default_rng(1).random((500, 7)) < 0.3).astype(np.uint8)
print(s.reduce(fl, np.arange(500) % 3, np.arange(500) % 41, 3, 41)[1])
big = ts.gen_corpus(9, 9000, 0, 4096, n_groups=2)           # 37 MB through the host path: two slabs, classified slab by slab
s2 = ts.Scanner(0, int(big.off[-1]) + 4096, big.n_files, 2)
r5 = s2.scan(big, 0, reuse=True)
print("streamed", r5["totals"], int(r5["global_counts"].sum()))
print("totals", r["totals"], r2["totals"], r3["totals"])
