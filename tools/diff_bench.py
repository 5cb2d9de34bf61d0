#!/usr/bin/env python3
"""BASELINE config C5 on one GPU: N (old, new) revision pairs through tsm_diff_pairs.

    python tools/diff_bench.py [pairs=50000] [cap=65536]

old ~ the C4 size law capped at `cap` bytes; new = old with Poisson(6) line edits.
Checks size-independent invariants (and a sampled oracle comparison), prints pairs/s end to end
(H2D + line hashes + Myers + D2H).
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tosem-2021-replication_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import tosemscan as ts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
t0 = time.time()
base = ts.gen_corpus(0x7053454D0005, n, size_law=1, pinned=False)
olds = [base.file_bytes(i)[:cap] for i in range(n)]
olds = [o[:o.rfind(b"\n") + 1] if len(o) == cap else o for o in olds]          # cut at a line end
news = [ts.gen_edit(0x7053454D0005 + i, o, 6.0) for i, o in enumerate(olds)]
a, b = ts.pack(olds, [1] * n, pinned=True), ts.pack(news, [1] * n, pinned=True)
print("generated %d pairs, %.1f + %.1f MB in %.1f s" % (n, a.source_bytes / 1e6, b.source_bytes / 1e6, time.time() - t0))
sc = ts.Scanner(0, 1 << 20, 16, 1)
sc.diff_pairs(a, b)
t0 = time.time()
add, rem = sc.diff_pairs(a, b)
dt = time.time() - t0
nl_old = np.array([o.count(b"\n") + (1 if o and not o.endswith(b"\n") else 0) for o in olds])
nl_new = np.array([o.count(b"\n") + (1 if o and not o.endswith(b"\n") else 0) for o in news])
assert np.array_equal(add - rem, nl_new - nl_old), "added - removed must equal the change in line count"
assert (add >= 0).all() and (rem >= 0).all() and (rem <= nl_old).all() and (add <= nl_new).all()
same = np.array([o == m for o, m in zip(olds, news)])
assert (add[same] == 0).all() and (rem[same] == 0).all()
import orc
idx = np.arange(0, n, max(1, n // 300))
sa = ts.pack([olds[i] for i in idx], [1] * len(idx))
sb = ts.pack([news[i] for i in idx], [1] * len(idx))
wa, wr = orc.diff_pairs((sa.arena, sa.off, sa.len), (sb.arena, sb.off, sb.len))
assert np.array_equal(add[idx], wa) and np.array_equal(rem[idx], wr), "sampled pairs differ from the oracle"
print("C5: %d pairs in %.3f s -> %.0f pairs/s, %.1f MB/s of revision text; cloc=%d added=%d removed=%d; %d sampled pairs match the oracle"
      % (n, dt, n / dt, (a.source_bytes + b.source_bytes) / dt / 1e6, int(add.sum() + rem.sum()), int(add.sum()), int(rem.sum()), len(idx)))
