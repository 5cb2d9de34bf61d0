"""CPU tests of the C-ABI library and the host-side logic (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orc
import tosemscan as ts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tosemscan.h")).read()
    declared = sorted(set(re.findall(r"\b(tsm_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(ts.SYMBOLS)
    L = ts.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.tsm_abi_version() == 1


def test_category_table_is_the_same_on_both_sides():
    for i in range(128):
        a, b = ts.category_name(i), orc.category_name(i)
        assert a == b, (i, a, b)
    assert ts.category_name(1) == "assertEqual" and ts.category_name(127) == "<other>"


def test_struct_layouts():
    assert ts.FILE_STAT.itemsize == 24 and ts.ASSERT_EVENT.itemsize == 32 and ts.HEADER_EVENT.itemsize == 16
    assert ts.FILE_STAT == orc.FILE_STAT and ts.ASSERT_EVENT == orc.ASSERT_EVENT and ts.HEADER_EVENT == orc.HEADER_EVENT


def test_no_cpu_fallback():
    from conftest import _have_gpu
    if _have_gpu():
        pytest.skip("GPU present")
    with pytest.raises(ts.TsmError) as e:
        ts.Scanner(device=0)
    assert e.value.status == -4      # TSM_E_CUDA: the product path fails loudly without a device


def test_layout_and_pack():
    length = np.array([0, 1, 128, 129, 4096], np.int32)
    off = np.zeros(6, np.int32)
    total = ts.lib().tsm_layout(length.ctypes.data_as(C.c_void_p), 5, off.ctypes.data_as(C.c_void_p))
    assert off.tolist() == [0, 0, 128, 256, 512, 4608] and total == 4608
    c = ts.pack([b"", b"a", b"b" * 128, b"c" * 129], [1, 2, 3, 4])
    a2, off2, len2 = orc.pack([b"", b"a", b"b" * 128, b"c" * 129])
    assert np.array_equal(c.off, off2) and np.array_equal(c.len, len2) and np.array_equal(c.arena[:off2[-1]], a2[:off2[-1]])
    assert c.file_bytes(3) == b"c" * 129 and c.algorithmic_bytes == 258 + 4 * 5


def test_synthetic_corpus_is_deterministic_and_has_the_named_shape():
    a = ts.gen_corpus(0x7053454D0002, 300, size_law=0, fixed_size=4096, n_groups=9, pinned=False)
    b = ts.gen_corpus(0x7053454D0002, 300, size_law=0, fixed_size=4096, n_groups=9, pinned=False)
    assert np.array_equal(a.arena, b.arena) and np.array_equal(a.ext, b.ext) and np.array_equal(a.grp, b.grp)
    assert (a.len == 4096).all() and a.source_bytes == 300 * 4096
    # sharding: slot i of the (first=1, stride=2) shard is logical file 1 + 2i
    s = ts.gen_corpus(0x7053454D0002, 150, 0, 4096, first_index=1, index_stride=2, n_groups=9, pinned=False)
    for i in (0, 7, 149):
        assert s.file_bytes(i) == a.file_bytes(1 + 2 * i) and s.ext[i] == a.ext[1 + 2 * i]
    res = orc.scan(a.arena, a.off, a.len, a.ext, a.grp, 9, events=False)
    lines = int(res["stats"]["n_lines"].sum())
    asserts = int(res["stats"]["n_assert"].sum())
    assert 30 < a.source_bytes / lines < 42            # ~36 B / line
    assert 0.06 < asserts / lines < 0.12               # ~8.8 % assertion lines
    exts = np.bincount(a.ext, minlength=7) / 300
    assert 0.35 < exts[1] < 0.65 and 0.25 < exts[2] < 0.55 and 0.02 < exts[4] < 0.2
    # every file ends with a newline and every byte is printable ASCII, CR or LF
    assert all(a.file_bytes(i).endswith(b"\n") for i in range(300))


def test_zipf_sizes_and_edits():
    z = ts.gen_corpus(0x7053454D0004, 400, size_law=1, pinned=False)
    assert z.len.min() >= 128 and z.len.max() <= (1 << 20) + 256
    assert np.median(z.len) < 2000 < z.len.mean() * 4          # heavy tail
    src = z.file_bytes(int(np.argmax(z.len < 8000)))
    e1, e2 = ts.gen_edit(5, src, 6.0), ts.gen_edit(5, src, 6.0)
    assert e1 == e2 and e1 != src
    assert ts.gen_edit(9, src, 0.0) == src                     # lambda 0: no edits


def test_two_rank_shards_cover_the_corpus_gloo(tmp_path):
    """N>1 host logic on CPU (gloo, world_size 2): round-robin shards + the one allreduce of counts."""
    import subprocess
    import sys
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import tosemscan as ts, orc
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
n = 64
c = ts.gen_corpus(77, n // w, 0, 2048, first_index=r, index_stride=w, n_groups=3, pinned=False)
res = orc.scan(c.arena, c.off, c.len, c.ext, c.grp, 3, events=False)   # checker stands in for the GPU scan
t = torch.from_numpy(np.concatenate([res["group_counts"].ravel(), res["global_counts"]]))
dist.all_reduce(t)
if r == 0:
    full = ts.gen_corpus(77, n, 0, 2048, n_groups=3, pinned=False)
    want = orc.scan(full.arena, full.off, full.len, full.ext, full.grp, 3, events=False)
    assert np.array_equal(t.numpy(), np.concatenate([want["group_counts"].ravel(), want["global_counts"]]))
    print("OK")
dist.destroy_process_group()
''' % (os.path.join(ROOT, "tests"), os.path.join(ROOT, "tosem-2021-replication_b200")))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-3000:]
