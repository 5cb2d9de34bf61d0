#!/usr/bin/env python3
"""BASELINE config C5 on one GPU, a few resident steps (for ncu): python tools/diff_prof.py [pairs=50000] [steps=2]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tosem-2021-replication_b200"))
import tosemscan as ts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
a, b = ts.gen_pairs(0x7053454D0005, n)
sc = ts.Scanner(0, 1 << 20, 16, 1)
sc.diff_upload(a, b)
for _ in range(steps):
    add, rem, det = sc.diff_resident(True)
print("C5: %d pairs, %.1f MB, kernels(ms) scan/myers/trace = %s, cloc=%d" % (n, (a.source_bytes + b.source_bytes) / 1e6,
      ["%.3f" % m for m in sc.diff_last_ms()], int(add.sum() + rem.sum())))
