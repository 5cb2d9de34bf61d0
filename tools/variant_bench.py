#!/usr/bin/env python3
"""Time the scan kernels of one build variant (TOSEMSCAN_LIB=... python tools/variant_bench.py [files])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tosem-2021-replication_b200"))
import numpy as np
import tosemscan as ts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
law = int(sys.argv[2]) if len(sys.argv) > 2 else 0
c = ts.gen_corpus(0x7053454D0002 if law == 0 else 0x7053454D0004, n, law, 4096, n_groups=9)
sc = ts.Scanner(0, int(c.off[-1]) + 4096, n, 16)
sc.upload(c)
for _ in range(3):
    sc.scan_resident(0)
sc.kernel_ms_stats(reset=True)
for _ in range(20):
    sc.scan_resident(0)
sums, k = sc.kernel_ms_stats(reset=True)
res = sc.download(0)
chk = int(np.bitwise_xor.reduce(res["stats"]["digest"])) ^ int(res["global_counts"].sum()) ^ int(res["totals"].sum())
print("%-28s plan %.4f scan %.4f classify %.4f totals %.4f ms | GB/s(scan) %.0f | check %016x" % (
    os.path.basename(os.environ.get("TOSEMSCAN_LIB", "default")), sums[0] / k, sums[1] / k, sums[2] / k, sums[3] / k,
    c.algorithmic_bytes / (sums[1] / k * 1e-3) / 1e9, chk))
