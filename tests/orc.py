"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load
this; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")
K = 128

FILE_STAT = np.dtype([("n_lines", "<u4"), ("n_assert", "<u4"), ("n_headers", "<u4"),
                      ("n_fixture", "<u4"), ("digest", "<u8")])
ASSERT_EVENT = np.dtype([("file", "<u4"), ("line_off", "<u4"), ("stmt_off", "<u4"),
                         ("stmt_len", "<u2"), ("cat", "<u2"), ("ident_off", "<u4"),
                         ("ident_len", "<u2"), ("pad", "<u2"), ("stmt_hash", "<u8")])
HEADER_EVENT = np.dtype([("file", "<u4"), ("line_off", "<u4"), ("line_len", "<u4"), ("kind", "<u4")])
DIFF_DETAIL = np.dtype([("hunks_add", "<i8"), ("hunks_del", "<i8"), ("hunks_mod", "<i8"),
                        ("added_assert", "<i8"), ("removed_assert", "<i8")])
assert FILE_STAT.itemsize == 24 and ASSERT_EVENT.itemsize == 32 and HEADER_EVENT.itemsize == 16

_lib = None


def build():
    so = os.path.join(ORC_DIR, "liborc.so")
    srcs = [os.path.join(ORC_DIR, f) for f in ("orc.c", "orc_mt.c", "orc.h", "orc_categories.inc")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORC_DIR, "-s"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_bytes_hash.restype = C.c_uint64
        L.orc_bytes_hash.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_line_hash.restype = C.c_uint64
        L.orc_line_hash.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_classify.restype = C.c_int
        L.orc_classify.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_category_name.restype = C.c_char_p
        L.orc_category_name.argtypes = [C.c_int]
        L.orc_statement.restype = None
        L.orc_statement.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_is_assert_line.restype = C.c_int
        L.orc_is_assert_line.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_header_kind.restype = C.c_int
        L.orc_header_kind.argtypes = [C.c_int, C.c_void_p, C.c_uint32]
        L.orc_method_string.restype = C.c_uint32
        L.orc_method_string.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_scan.restype = C.c_int
        L.orc_scan.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int32] + [C.c_void_p] * 3 + \
            [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_scan_ex.restype = C.c_int
        L.orc_scan_ex.argtypes = L.orc_scan.argtypes + [C.c_uint32]
        L.orc_lcs.restype = C.c_int64
        L.orc_lcs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.orc_diff_pairs.restype = C.c_int
        L.orc_diff_pairs.argtypes = [C.c_void_p] * 6 + [C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_diff_pairs_detail.restype = C.c_int
        L.orc_diff_pairs_detail.argtypes = [C.c_void_p] * 8 + [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_diff_script.restype = C.c_int64
        L.orc_diff_script.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_statements.restype = C.c_int64
        L.orc_statements.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_reduce.restype = C.c_int
        L.orc_reduce.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p]
        L.orc_ngram_hashes.restype = None
        L.orc_ngram_hashes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_mt_affinity_cpus.restype = C.c_int
        L.orc_mt_create.restype = C.c_int
        L.orc_mt_create.argtypes = [C.c_int, C.c_int32]
        L.orc_mt_destroy.restype = None
        L.orc_mt_scan.restype = C.c_int
        L.orc_mt_scan.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int32] + [C.c_void_p] * 3
        L.orc_mt_diff.restype = C.c_int
        L.orc_mt_diff.argtypes = [C.c_void_p] * 8 + [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bytes_hash(b: bytes) -> int:
    buf = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8)
    return int(lib().orc_bytes_hash(_p(buf), len(b)))


def line_hash(b: bytes) -> int:
    buf = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8)
    return int(lib().orc_line_hash(_p(buf), len(b)))


def category_name(i: int) -> str:
    return lib().orc_category_name(i).decode()


def classify(t: bytes):
    buf = np.frombuffer(t, dtype=np.uint8) if len(t) else np.zeros(1, np.uint8)
    io, il = C.c_uint32(), C.c_uint32()
    cat = lib().orc_classify(_p(buf), len(t), C.byref(io), C.byref(il))
    return cat, io.value, il.value


def category_string(t: bytes) -> str:
    """Category cell as the lost tool printed it (verbatim identifier for OTHER)."""
    cat, io, il = classify(t)
    if cat == 127:
        return t[io:io + il].decode("latin-1")
    return category_name(cat)


def statement(line: bytes) -> bytes:
    buf = np.frombuffer(line, dtype=np.uint8) if len(line) else np.zeros(1, np.uint8)
    so, sl = C.c_uint32(), C.c_uint32()
    lib().orc_statement(_p(buf), len(line), C.byref(so), C.byref(sl))
    return line[so.value:so.value + sl.value]


def is_assert_line(line: bytes) -> bool:
    buf = np.frombuffer(line, dtype=np.uint8) if len(line) else np.zeros(1, np.uint8)
    return bool(lib().orc_is_assert_line(_p(buf), len(line)))


def header_kind(ext: int, line: bytes) -> int:
    buf = np.frombuffer(line, dtype=np.uint8) if len(line) else np.zeros(1, np.uint8)
    return int(lib().orc_header_kind(ext, _p(buf), len(line)))


def method_string(ext: int, line: bytes) -> bytes:
    buf = np.frombuffer(line, dtype=np.uint8) if len(line) else np.zeros(1, np.uint8)
    out = np.zeros(len(line) + 8, np.uint8)
    n = lib().orc_method_string(ext, _p(buf), len(line), _p(out), out.size)
    return out[:n].tobytes()


def scan(arena, off, length, ext, grp, n_groups=1, events=True, line_hashes=False, rev_b=False):
    """Run the oracle over a packed corpus.  Returns a dict of numpy arrays."""
    n = len(length)
    arena = np.ascontiguousarray(arena, np.uint8)
    off = np.ascontiguousarray(off, np.int32)
    length = np.ascontiguousarray(length, np.int32)
    ext = np.ascontiguousarray(ext, np.uint8)
    grp = np.ascontiguousarray(grp, np.uint16)
    stats = np.zeros(n, FILE_STAT)
    gc = np.zeros((n_groups, K), np.int64)
    glob = np.zeros(K, np.int64)
    na, nh = C.c_int64(), C.c_int64()
    L = lib()
    # first pass counts events, second fills
    fl = 8 if rev_b else 0
    rc = L.orc_scan_ex(_p(arena), _p(off), _p(length), _p(ext), _p(grp), n, n_groups, _p(stats), _p(gc),
                       _p(glob), None, 0, C.byref(na), None, 0, C.byref(nh), None, None, fl)
    if rc != 0:
        raise ValueError("orc_scan: malformed corpus")
    out = {"stats": stats, "group_counts": gc, "global_counts": glob}
    if events or line_hashes:
        aev = np.zeros(max(na.value, 1), ASSERT_EVENT)
        hev = np.zeros(max(nh.value, 1), HEADER_EVENT)
        nl = int(stats["n_lines"].astype(np.int64).sum())
        lh = np.zeros(max(nl, 1), np.uint64) if line_hashes else None
        lb = np.zeros(n + 1, np.int64) if line_hashes else None
        L.orc_scan_ex(_p(arena), _p(off), _p(length), _p(ext), _p(grp), n, n_groups, _p(stats), _p(gc),
                      _p(glob), _p(aev), aev.size, C.byref(na), _p(hev), hev.size, C.byref(nh), _p(lh), _p(lb), fl)
        out["assert_events"] = aev[:na.value]
        out["header_events"] = hev[:nh.value]
        if line_hashes:
            out["line_hash"] = lh[:nl]
            out["line_base"] = lb
    return out


def ngram_hashes(line_hash, line_base, n):
    line_hash = np.ascontiguousarray(line_hash, np.uint64)
    line_base = np.ascontiguousarray(line_base, np.int64)
    out = np.zeros(max(len(line_hash), 1), np.uint64)
    lib().orc_ngram_hashes(_p(line_hash), _p(line_base), len(line_base) - 1, int(n), _p(out))
    return out[:len(line_hash)]


def line_records(arena, off, length, ext):
    """(line_base, line_hash, line_end, line_flag) of every line, files in order: the oracle's side of tsm_line_hashes."""
    n = len(length)
    res = scan(arena, off, length, ext, np.zeros(n, np.uint16), 1, events=True, line_hashes=True)
    base, lh = res["line_base"], res["line_hash"]
    end = np.zeros(len(lh), np.uint32)
    flag = np.zeros(len(lh), np.uint8)
    starts = np.zeros(len(lh), np.int64)
    for f in range(n):
        b = arena[int(off[f]):int(off[f]) + int(length[f])]
        nl = np.nonzero(b == 10)[0]
        k = int(base[f + 1] - base[f])
        e = np.full(k, int(length[f]), np.int64)
        e[:len(nl)] = nl[:k] if len(nl) >= k else nl
        end[int(base[f]):int(base[f + 1])] = e
        st = np.zeros(k, np.int64)
        st[1:] = e[:-1] + 1
        starts[int(base[f]):int(base[f + 1])] = st
    ev = res["assert_events"]
    if len(ev):
        key = {(int(f), int(o)) for f, o in zip(ev["file"], ev["line_off"])}
        for f in range(n):
            for i in range(int(base[f]), int(base[f + 1])):
                if (f, int(starts[i])) in key:
                    flag[i] = 1
    return base, lh, end, flag


class MtScanner:
    """orc_scan over a pool of persistent POSIX threads (oracle/orc_mt.c): the host-cores baseline.  Buffers
    are allocated once; scan() is one C call with no Python work per file or per thread."""

    def __init__(self, threads=0, max_groups=16):
        self.threads = int(lib().orc_mt_create(int(threads), int(max_groups)))
        if self.threads < 1:
            raise RuntimeError("orc_mt_create failed")
        self._bufs = None

    def scan(self, arena, off, length, ext, grp, n_groups=1):
        n = len(length)
        if self._bufs is None or self._bufs[0].size != n or self._bufs[1].shape[0] != n_groups:
            self._bufs = (np.zeros(n, FILE_STAT), np.zeros((n_groups, K), np.int64), np.zeros(K, np.int64))
        stats, gc, glob = self._bufs
        rc = lib().orc_mt_scan(_p(arena), _p(off), _p(length), _p(ext), _p(grp), n, n_groups, _p(stats), _p(gc), _p(glob))
        if rc != 0:
            raise ValueError("orc_mt_scan failed (%d)" % rc)
        return {"stats": stats, "group_counts": gc, "global_counts": glob}

    def diff(self, a, b):
        """(added, removed, detail) of the pairs (a[i], b[i]); a, b = (arena, off, len, ext)."""
        n = len(a[2])
        if getattr(self, "_dbufs", None) is None or self._dbufs[0].size != n:
            self._dbufs = (np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(max(n, 1), DIFF_DETAIL))
        add, rem, det = self._dbufs
        rc = lib().orc_mt_diff(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(b[0]), _p(b[1]), _p(b[2]), _p(b[3]), n, _p(add), _p(rem), _p(det))
        if rc != 0:
            raise ValueError("orc_mt_diff failed (%d)" % rc)
        return add, rem, det[:n]

    def close(self):
        lib().orc_mt_destroy()


def lcs(a, b) -> int:
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    return int(lib().orc_lcs(_p(a), a.size, _p(b), b.size))


def diff_pairs(old, new):
    """old/new: (arena, off, len) triples with equal file counts."""
    n = len(old[2])
    added = np.zeros(n, np.int64)
    removed = np.zeros(n, np.int64)
    a = [np.ascontiguousarray(old[0], np.uint8), np.ascontiguousarray(old[1], np.int32),
         np.ascontiguousarray(old[2], np.int32)]
    b = [np.ascontiguousarray(new[0], np.uint8), np.ascontiguousarray(new[1], np.int32),
         np.ascontiguousarray(new[2], np.int32)]
    rc = lib().orc_diff_pairs(_p(a[0]), _p(a[1]), _p(a[2]), _p(b[0]), _p(b[1]), _p(b[2]), n,
                              _p(added), _p(removed))
    if rc != 0:
        raise ValueError("orc_diff_pairs failed")
    return added, removed


def diff_pairs_detail(old, new):
    """old/new: (arena, off, len, ext).  Returns added, removed, detail (DIFF_DETAIL records)."""
    n = len(old[2])
    added = np.zeros(n, np.int64)
    removed = np.zeros(n, np.int64)
    det = np.zeros(max(n, 1), DIFF_DETAIL)
    a = [np.ascontiguousarray(old[0], np.uint8), np.ascontiguousarray(old[1], np.int32),
         np.ascontiguousarray(old[2], np.int32), np.ascontiguousarray(old[3], np.uint8)]
    b = [np.ascontiguousarray(new[0], np.uint8), np.ascontiguousarray(new[1], np.int32),
         np.ascontiguousarray(new[2], np.int32), np.ascontiguousarray(new[3], np.uint8)]
    rc = lib().orc_diff_pairs_detail(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(b[0]), _p(b[1]), _p(b[2]), _p(b[3]),
                                     n, _p(added), _p(removed), _p(det))
    if rc != 0:
        raise ValueError("orc_diff_pairs_detail failed")
    return added, removed, det[:n]


def diff_script(a, b, fa=None, fb=None):
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    det = np.zeros(1, DIFF_DETAIL)
    D = lib().orc_diff_script(_p(a), a.size, _p(b), b.size, None if fa is None else _p(np.ascontiguousarray(fa, np.uint8)),
                              None if fb is None else _p(np.ascontiguousarray(fb, np.uint8)), _p(det))
    return int(D), det[0]


def statements(arena, off, length):
    """SPEC section 10: (line_base, line_end, line_kind) of a packed corpus."""
    arena = np.ascontiguousarray(arena, np.uint8)
    off = np.ascontiguousarray(off, np.int32)
    length = np.ascontiguousarray(length, np.int32)
    n = len(length)
    base = np.zeros(n + 1, np.int64)
    total = lib().orc_statements(_p(arena), _p(off), _p(length), n, _p(base), None, None, 0)
    end = np.zeros(max(total, 1), np.uint32)
    kind = np.zeros(max(total, 1), np.uint8)
    lib().orc_statements(_p(arena), _p(off), _p(length), n, _p(base), _p(end), _p(kind), total)
    return base, end[:total], kind[:total]


def statement_texts(data: bytes):
    """Statements of one file as the lost tool printed them: stripped lines joined with one blank."""
    arena, off, length = pack([data])
    base, end, kind = statements(arena, off, length)
    out, cur, pos = [], None, 0
    for e, k in zip(end.tolist(), kind.tolist()):
        line = data[pos:e].strip(b" \t\r\x0b\x0c")
        if k == 1:
            if cur is not None:
                out.append(b" ".join(cur))
            cur = [line]
        elif k == 2:
            cur.append(line)
        pos = e + 1
    if cur is not None:
        out.append(b" ".join(cur))
    return out


def reduce(flags, repo, case_id, n_repos, n_cases):
    flags = np.ascontiguousarray(flags, np.uint8)
    repo = np.ascontiguousarray(repo, np.int32)
    case_id = np.ascontiguousarray(case_id, np.int32)
    n_rows, n_flags = flags.shape
    out = np.zeros((n_flags, n_repos), np.int64)
    cpr = np.zeros(n_repos, np.int64)
    rc = lib().orc_reduce(_p(flags), _p(repo), _p(case_id), n_rows, n_flags, n_repos, n_cases, _p(out), _p(cpr))
    if rc != 0:
        raise ValueError("orc_reduce failed")
    return out, cpr


def pack(files, align=128):
    """Pack a list of bytes objects into (arena, off, len) per docs/SPEC.md section 1."""
    n = len(files)
    off = np.zeros(n + 1, np.int32)
    length = np.zeros(n, np.int32)
    o = 0
    for i, f in enumerate(files):
        off[i] = o
        length[i] = len(f)
        o += (len(f) + align - 1) // align * align
    off[n] = o
    arena = np.zeros(max(o, align), np.uint8)
    for i, f in enumerate(files):
        if f:
            arena[off[i]:off[i] + len(f)] = np.frombuffer(f, np.uint8)
    return arena, off, length
