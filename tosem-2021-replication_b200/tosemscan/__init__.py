"""tosemscan - Python host mirror of include/tosemscan.h (ctypes over the C ABI).

The reference package ships no operator/plugin interface; this module is the thin
host layer the parity tests and bench.py use: numpy arrays in, numpy arrays out, every call
forwarded to ``libtosemscan.so`` (hand-written sm_100a CUDA).  There is NO CPU fallback: a
missing library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TOSEMSCAN_LIB") or os.path.join(_HERE, "libtosemscan.so")   # env override: tuning variants

K = 128
ALIGN = 128
EXT = {"py": 1, "cc": 2, "cpp": 3, "java": 4, "c": 5, "h": 6}
SCAN_ASSERT_EVENTS = 1
SCAN_HEADER_EVENTS = 2
SCAN_REV_B = 8            # docs/SPEC.md section 4b (golden G1)

FILE_STAT = np.dtype([("n_lines", "<u4"), ("n_assert", "<u4"), ("n_headers", "<u4"),
                      ("n_fixture", "<u4"), ("digest", "<u8")])
ASSERT_EVENT = np.dtype([("file", "<u4"), ("line_off", "<u4"), ("stmt_off", "<u4"),
                         ("stmt_len", "<u2"), ("cat", "<u2"), ("ident_off", "<u4"),
                         ("ident_len", "<u2"), ("pad", "<u2"), ("stmt_hash", "<u8")])
HEADER_EVENT = np.dtype([("file", "<u4"), ("line_off", "<u4"), ("line_len", "<u4"), ("kind", "<u4")])
DIFF_DETAIL = np.dtype([("hunks_add", "<i8"), ("hunks_del", "<i8"), ("hunks_mod", "<i8"),
                        ("added_assert", "<i8"), ("removed_assert", "<i8")])

# every symbol include/tosemscan.h declares (tests check the library exports exactly these)
SYMBOLS = ["tsm_abi_version", "tsm_strerror", "tsm_category_name", "tsm_create", "tsm_destroy", "tsm_scan",
           "tsm_upload", "tsm_scan_resident", "tsm_download", "tsm_device_counts", "tsm_last_launch_count", "tsm_last_kernel_ms", "tsm_kernel_ms_stats",
           "tsm_diff_pairs", "tsm_diff_pairs_detail", "tsm_statements", "tsm_line_hashes", "tsm_diff_upload", "tsm_diff_resident", "tsm_diff_last_ms", "tsm_reduce", "tsm_host_alloc", "tsm_host_free", "tsm_layout", "tsm_gen_sizes",
           "tsm_gen_fill", "tsm_gen_edit", "tsm_gen_pair_sizes", "tsm_gen_pair_fill"]


class TsmError(RuntimeError):
    def __init__(self, status, what):
        self.status = status
        super().__init__(f"{what}: {lib().tsm_strerror(status).decode()} ({status})")


class _Corpus(C.Structure):
    _fields_ = [("arena", C.c_void_p), ("off", C.c_void_p), ("len", C.c_void_p), ("ext", C.c_void_p),
                ("grp", C.c_void_p), ("n_files", C.c_int32), ("n_groups", C.c_int32)]


class _Result(C.Structure):
    _fields_ = [("stats", C.c_void_p), ("group_counts", C.c_void_p), ("global_counts", C.c_void_p),
                ("aev", C.c_void_p), ("aev_cap", C.c_int64), ("n_aev", C.c_int64),
                ("hev", C.c_void_p), ("hev_cap", C.c_int64), ("n_hev", C.c_int64),
                ("totals", C.c_int64 * 4)]


_lib = None


def lib():
    """Load libtosemscan.so (built in-tree by __graft_entry__.build() / make). Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C tosem-2021-replication_b200` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.tsm_abi_version.restype = C.c_int
        L.tsm_strerror.restype = C.c_char_p
        L.tsm_strerror.argtypes = [C.c_int]
        L.tsm_category_name.restype = C.c_char_p
        L.tsm_category_name.argtypes = [C.c_int]
        L.tsm_create.restype = C.c_int
        L.tsm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int32, C.c_int32, C.c_int64]
        L.tsm_destroy.restype = None
        L.tsm_destroy.argtypes = [C.c_void_p]
        L.tsm_scan.restype = C.c_int
        L.tsm_scan.argtypes = [C.c_void_p, C.POINTER(_Corpus), C.POINTER(_Result), C.c_uint32, C.c_void_p]
        L.tsm_upload.restype = C.c_int
        L.tsm_upload.argtypes = [C.c_void_p, C.POINTER(_Corpus), C.c_void_p]
        L.tsm_scan_resident.restype = C.c_int
        L.tsm_scan_resident.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.tsm_download.restype = C.c_int
        L.tsm_download.argtypes = [C.c_void_p, C.POINTER(_Result), C.c_void_p]
        L.tsm_device_counts.restype = C.c_int
        L.tsm_device_counts.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.tsm_last_launch_count.restype = C.c_int
        L.tsm_last_launch_count.argtypes = [C.c_void_p]
        L.tsm_last_kernel_ms.restype = C.c_int
        L.tsm_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float * 4)]
        L.tsm_kernel_ms_stats.restype = C.c_int
        L.tsm_kernel_ms_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double * 4), C.POINTER(C.c_int64), C.c_int]
        L.tsm_diff_pairs.restype = C.c_int
        L.tsm_diff_pairs.argtypes = [C.c_void_p, C.POINTER(_Corpus), C.POINTER(_Corpus), C.c_void_p, C.c_void_p, C.c_void_p]
        L.tsm_diff_pairs_detail.restype = C.c_int
        L.tsm_diff_pairs_detail.argtypes = [C.c_void_p, C.POINTER(_Corpus), C.POINTER(_Corpus), C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]
        L.tsm_statements.restype = C.c_int
        L.tsm_statements.argtypes = [C.c_void_p, C.POINTER(_Corpus), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.POINTER(C.c_int64), C.c_void_p]
        L.tsm_line_hashes.restype = C.c_int
        L.tsm_line_hashes.argtypes = [C.c_void_p, C.POINTER(_Corpus)] + [C.c_void_p] * 4 + [C.c_int64, C.POINTER(C.c_int64), C.c_int32,
                                      C.c_void_p, C.c_void_p]
        L.tsm_reduce.restype = C.c_int
        L.tsm_reduce.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 4 + [C.c_void_p] * 3
        L.tsm_host_alloc.restype = C.c_void_p
        L.tsm_host_alloc.argtypes = [C.c_int64]
        L.tsm_host_free.restype = None
        L.tsm_host_free.argtypes = [C.c_void_p]
        L.tsm_layout.restype = C.c_int64
        L.tsm_layout.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.tsm_gen_sizes.restype = C.c_int
        L.tsm_gen_sizes.argtypes = [C.c_uint64, C.c_int32, C.c_int, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.tsm_gen_fill.restype = C.c_int
        L.tsm_gen_fill.argtypes = [C.c_uint64, C.c_int32, C.c_int, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tsm_gen_edit.restype = C.c_int64
        L.tsm_gen_edit.argtypes = [C.c_uint64, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int64]
        L.tsm_gen_pair_sizes.restype = C.c_int
        L.tsm_gen_pair_sizes.argtypes = [C.c_uint64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.tsm_gen_pair_fill.restype = C.c_int
        L.tsm_gen_pair_fill.argtypes = [C.c_uint64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double] + [C.c_void_p] * 7
        L.tsm_diff_upload.restype = C.c_int
        L.tsm_diff_upload.argtypes = [C.c_void_p, C.POINTER(_Corpus), C.POINTER(_Corpus), C.c_void_p]
        L.tsm_diff_resident.restype = C.c_int
        L.tsm_diff_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tsm_diff_last_ms.restype = C.c_int
        L.tsm_diff_last_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float * 3)]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def category_name(i):
    return lib().tsm_category_name(i).decode()


class _Pinned:
    """A pinned host allocation (cudaHostAlloc) exposed as a numpy uint8 array."""

    def __init__(self, nbytes):
        self.ptr = lib().tsm_host_alloc(nbytes)
        self.nbytes = nbytes
        if not self.ptr:
            raise MemoryError("tsm_host_alloc failed")
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().tsm_host_free(self.ptr)
            self.ptr = None


def host_buffer(nbytes, pinned=True):
    """uint8 buffer for an arena: pinned when a GPU is present and pinned=True, else plain numpy."""
    if pinned:
        try:
            pin = _Pinned(max(int(nbytes), ALIGN))
            arr = pin.array
            arr[:] = 0
            return arr, pin
        except MemoryError:
            pass
    return np.zeros(max(int(nbytes), ALIGN), np.uint8), None


class Corpus:
    """Packed corpus of docs/SPEC.md section 1 (arena + int32 offset index + per-file tags)."""

    def __init__(self, arena, off, length, ext, grp=None, n_groups=1, keep=None):
        self.arena = arena
        self.off = np.ascontiguousarray(off, np.int32)
        self.len = np.ascontiguousarray(length, np.int32)
        self.ext = np.ascontiguousarray(ext, np.uint8)
        self.grp = np.zeros(len(self.len), np.uint16) if grp is None else np.ascontiguousarray(grp, np.uint16)
        self.n_groups = int(n_groups)
        self._keep = keep
        if keep is not None:                                 # pinned arena: pin the index too (tsm_scan copies it every call;
            self._pin_index()                               # from pageable memory those copies are staged and block the host)

    def _pin_index(self):
        pins = []
        for name in ("off", "len", "ext", "grp"):
            a = getattr(self, name)
            try:
                pin = _Pinned(max(a.nbytes, ALIGN))
            except MemoryError:
                return
            b = pin.array[:a.nbytes].view(a.dtype)
            b[:] = a
            setattr(self, name, b)
            pins.append(pin)
        self._index_pins = pins

    @property
    def n_files(self):
        return len(self.len)

    @property
    def source_bytes(self):
        return int(self.len.astype(np.int64).sum())

    @property
    def algorithmic_bytes(self):
        """every source byte once + the int32 offset index."""
        return self.source_bytes + 4 * (self.n_files + 1)

    def c_struct(self):
        return _Corpus(_p(self.arena), _p(self.off), _p(self.len), _p(self.ext), _p(self.grp),
                       self.n_files, self.n_groups)

    def file_bytes(self, i):
        o = int(self.off[i])
        return self.arena[o:o + int(self.len[i])].tobytes()


def pack(files, exts, grps=None, n_groups=1, pinned=False):
    """Pack bytes objects end to end with 128-B aligned starts."""
    n = len(files)
    length = np.array([len(f) for f in files], np.int32)
    off = np.zeros(n + 1, np.int32)
    total = lib().tsm_layout(_p(length), n, _p(off))
    if total < 0:
        raise ValueError("corpus does not fit an int32-indexed arena (2 GiB): pack it in batches")
    arena, keep = host_buffer(total, pinned)
    for i, f in enumerate(files):
        if f:
            arena[off[i]:off[i] + len(f)] = np.frombuffer(f, np.uint8)
    return Corpus(arena, off, length, exts, grps, n_groups, keep)


def gen_corpus(seed, n_files, size_law=0, fixed_size=4096, first_index=0, index_stride=1, n_groups=1,
               pinned=True, threads=None):
    """Synthetic corpus of SURVEY.md section 8d (deterministic, std::mt19937_64 in the C++ host library)."""
    L = lib()
    length = np.zeros(n_files, np.int32)
    ext = np.zeros(n_files, np.uint8)
    grp = np.zeros(n_files, np.uint16)
    rc = L.tsm_gen_sizes(seed, n_files, size_law, fixed_size, first_index, index_stride, _p(length), _p(ext),
                         _p(grp), n_groups)
    if rc:
        raise TsmError(rc, "tsm_gen_sizes")
    off = np.zeros(n_files + 1, np.int32)
    total = L.tsm_layout(_p(length), n_files, _p(off))
    if total < 0:
        raise ValueError("corpus does not fit an int32-indexed arena")
    arena, keep = host_buffer(total, pinned)
    # files are independent: fill slices on all host cores (ctypes releases the GIL)
    nthr = max(1, min(threads or (os.cpu_count() or 1), 64, (n_files + 255) // 256))
    bounds = [n_files * t // nthr for t in range(nthr + 1)]

    def fill(t):
        a, b = bounds[t], bounds[t + 1]
        if a == b:
            return 0
        return L.tsm_gen_fill(seed, b - a, size_law, first_index + a * index_stride, index_stride,
                              _p(off[a:b + 1]), _p(length[a:b]), _p(ext[a:b]), _p(arena))
    if nthr == 1:
        rcs = [fill(0)]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nthr) as ex:
            rcs = list(ex.map(fill, range(nthr)))
    if any(rcs):
        raise TsmError([r for r in rcs if r][0], "tsm_gen_fill")
    return Corpus(arena, off, length, ext, grp, n_groups, keep)


def gen_pair_sizes(seed, n_pairs, cap=65536, lam=6.0, index=None, first_index=0, index_stride=1, threads=None):
    """(len_old, len_new, ext) of the logical pairs index[...] (or first_index + i*stride) of BASELINE config C5."""
    L = lib()
    idx = None if index is None else np.ascontiguousarray(index, np.int32)
    n = n_pairs if idx is None else len(idx)
    lo, ln, ext = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
    nthr = max(1, min(threads or (os.cpu_count() or 1), 64, (n + 255) // 256))
    bounds = [n * t // nthr for t in range(nthr + 1)]

    def sizes(t):
        a, b = bounds[t], bounds[t + 1]
        if a == b:
            return 0
        return L.tsm_gen_pair_sizes(seed, b - a, None if idx is None else _p(idx[a:b]), first_index + a * index_stride, index_stride,
                                    cap, float(lam), _p(lo[a:b]), _p(ln[a:b]), _p(ext[a:b]))
    _run_threads(sizes, nthr, "tsm_gen_pair_sizes")
    return lo, ln, ext


def _run_threads(fn, nthr, what):
    if nthr == 1:
        rcs = [fn(0)]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nthr) as ex:
            rcs = list(ex.map(fn, range(nthr)))
    if any(rcs):
        raise TsmError([r for r in rcs if r][0], what)


def gen_pairs(seed, n_pairs, cap=65536, lam=6.0, first_index=0, index_stride=1, pinned=True, threads=None, index=None, sizes=None):
    """BASELINE config C5: (olds, news) corpora of revision pairs, generated in C++.  `index` picks
    logical pairs by number (a rank's share of a size-balanced deal); `sizes` = gen_pair_sizes of the same pairs, if known."""
    L = lib()
    idx = None if index is None else np.ascontiguousarray(index, np.int32)
    n = n_pairs if idx is None else len(idx)
    lo, ln, ext = sizes if sizes is not None else gen_pair_sizes(seed, n, cap, lam, idx, first_index, index_stride, threads)
    oo, on = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
    to, tn = L.tsm_layout(_p(lo), n, _p(oo)), L.tsm_layout(_p(ln), n, _p(on))
    if to < 0 or tn < 0:
        raise ValueError("pairs do not fit an int32-indexed arena")
    ao, ko = host_buffer(to, pinned)
    an, kn = host_buffer(tn, pinned)
    nthr = max(1, min(threads or (os.cpu_count() or 1), 64, (n + 255) // 256))
    bounds = [n * t // nthr for t in range(nthr + 1)]

    def fill(t):
        a, b = bounds[t], bounds[t + 1]
        if a == b:
            return 0
        return L.tsm_gen_pair_fill(seed, b - a, None if idx is None else _p(idx[a:b]), first_index + a * index_stride, index_stride,
                                   cap, float(lam), _p(ext[a:b]), _p(oo[a:b + 1]), _p(lo[a:b]), _p(ao), _p(on[a:b + 1]), _p(ln[a:b]), _p(an))
    _run_threads(fill, nthr, "tsm_gen_pair_fill")
    return Corpus(ao, oo, lo, ext, None, 1, ko), Corpus(an, on, ln, ext.copy(), None, 1, kn)


def gen_edit(seed, src: bytes, lam=6.0) -> bytes:
    buf = np.frombuffer(src, np.uint8) if src else np.zeros(1, np.uint8)
    out = np.zeros(len(src) + 64 * 256 * int(lam * 4 + 16), np.uint8)
    n = lib().tsm_gen_edit(seed, _p(buf), len(src), float(lam), _p(out), out.size)
    if n < 0:
        raise ValueError("tsm_gen_edit failed")
    return out[:n].tobytes()


class Scanner:
    """One CUDA device context (tsm_ctx)."""

    def __init__(self, device=0, max_arena_bytes=1 << 26, max_files=1 << 16, max_groups=16, max_events=0):
        self._ctx = C.c_void_p()
        rc = lib().tsm_create(C.byref(self._ctx), device, max_arena_bytes, max_files, max_groups, max_events)
        if rc:
            raise TsmError(rc, "tsm_create")
        self.max_events = max_events if max_events else max_arena_bytes // 32 + max_files
        self._corpus = None

    def close(self):
        if self._ctx:
            lib().tsm_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _result(self, n_files, n_groups, flags, event_cap, reuse=False):
        if reuse:                                            # pinned buffers kept by the Scanner: D2H at the link rate, no
            key = (n_files, n_groups)                        # allocation per call; the arrays are valid until the next such call
            if getattr(self, "_res_key", None) != key:
                bufs = [host_buffer(max(n_files, 1) * FILE_STAT.itemsize), host_buffer(n_groups * K * 8), host_buffer(K * 8)]
                self._res_pins = [b[1] for b in bufs]
                self._res_bufs = (bufs[0][0][:n_files * FILE_STAT.itemsize].view(FILE_STAT),
                                  bufs[1][0][:n_groups * K * 8].view(np.int64).reshape(n_groups, K), bufs[2][0][:K * 8].view(np.int64))
                self._res_key = key
            res = {"stats": self._res_bufs[0], "group_counts": self._res_bufs[1], "global_counts": self._res_bufs[2]}
        else:
            res = {"stats": np.zeros(n_files, FILE_STAT), "group_counts": np.zeros((n_groups, K), np.int64),
                   "global_counts": np.zeros(K, np.int64)}
        r = _Result(_p(res["stats"]), _p(res["group_counts"]), _p(res["global_counts"]), None, 0, 0, None, 0, 0)
        if flags & SCAN_ASSERT_EVENTS:
            res["assert_events"] = np.zeros(max(event_cap, 1), ASSERT_EVENT)
            r.aev, r.aev_cap = _p(res["assert_events"]), event_cap
        if flags & SCAN_HEADER_EVENTS:
            res["header_events"] = np.zeros(max(event_cap, 1), HEADER_EVENT)
            r.hev, r.hev_cap = _p(res["header_events"]), event_cap
        return res, r

    @staticmethod
    def _finish(res, r, flags):
        res["totals"] = np.array(list(r.totals), np.int64)
        if flags & SCAN_ASSERT_EVENTS:
            res["assert_events"] = res["assert_events"][:r.n_aev]
        if flags & SCAN_HEADER_EVENTS:
            res["header_events"] = res["header_events"][:r.n_hev]
        return res

    def scan(self, corpus, flags=0, stream=None, event_cap=None, reuse=False):
        """End-to-end host path: H2D + kernels + D2H (tsm_scan).  reuse=True: the per-file records and count tables land in
        pinned buffers the Scanner keeps (valid until the next scan with reuse=True)."""
        want_ev = flags & (SCAN_ASSERT_EVENTS | SCAN_HEADER_EVENTS)
        cap = int(event_cap if event_cap is not None else (max(corpus.source_bytes // 8 + 16, 1024) if want_ev else 0))
        res, r = self._result(corpus.n_files, corpus.n_groups, flags, cap, reuse)
        cs = corpus.c_struct()
        rc = lib().tsm_scan(self._ctx, C.byref(cs), C.byref(r), flags, stream)
        if rc:
            raise TsmError(rc, "tsm_scan")
        self._corpus = corpus                               # tsm_scan leaves this corpus resident: download() reads its results
        return self._finish(res, r, flags)

    def upload(self, corpus, stream=None):
        cs = corpus.c_struct()
        rc = lib().tsm_upload(self._ctx, C.byref(cs), stream)
        if rc:
            raise TsmError(rc, "tsm_upload")
        self._corpus = corpus

    def scan_resident(self, flags=0, stream=None):
        rc = lib().tsm_scan_resident(self._ctx, flags, stream)
        if rc:
            raise TsmError(rc, "tsm_scan_resident")

    def download(self, flags=0, stream=None, event_cap=None):
        c = self._corpus
        cap = int(event_cap if event_cap is not None else max(c.source_bytes // 8 + 16, 1024))
        res, r = self._result(c.n_files, c.n_groups, flags, cap)
        rc = lib().tsm_download(self._ctx, C.byref(r), stream)
        if rc:
            raise TsmError(rc, "tsm_download")
        return self._finish(res, r, flags)

    def device_counts(self):
        """(device pointer, n_int64) of the [n_groups+1][K]+4 count table, for the one allreduce."""
        ptr, n = C.c_void_p(), C.c_int64()
        rc = lib().tsm_device_counts(self._ctx, C.byref(ptr), C.byref(n))
        if rc:
            raise TsmError(rc, "tsm_device_counts")
        return ptr.value, n.value

    def last_launch_count(self):
        return lib().tsm_last_launch_count(self._ctx)

    def last_kernel_ms(self):
        """Device time of k_plan, k_scan, k_classify of the last scan (4th slot is 0: k_totals is fused)."""
        ms = (C.c_float * 4)()
        rc = lib().tsm_last_kernel_ms(self._ctx, C.byref(ms))
        if rc:
            raise TsmError(rc, "tsm_last_kernel_ms")
        return [float(x) for x in ms]

    def kernel_ms_stats(self, reset=False):
        """(sum of ms per kernel [plan, scan, classify, totals], number of scans) since the last reset."""
        sums, n = (C.c_double * 4)(), C.c_int64()
        rc = lib().tsm_kernel_ms_stats(self._ctx, C.byref(sums), C.byref(n), int(reset))
        if rc:
            raise TsmError(rc, "tsm_kernel_ms_stats")
        return [float(x) for x in sums], int(n.value)

    def line_hashes(self, corpus, ngram=0, stream=None):
        """S9: (line_base[n+1], line_hash, line_end, line_flag[, ngram_hash]) of every line, files in order."""
        cs = corpus.c_struct()
        base = np.zeros(corpus.n_files + 1, np.int64)
        n = C.c_int64()
        rc = lib().tsm_line_hashes(self._ctx, C.byref(cs), _p(base), None, None, None, 0, C.byref(n), 0, None, stream)
        if rc not in (0, -3):
            raise TsmError(rc, "tsm_line_hashes")
        t = max(int(n.value), 1)
        lh, le, lf = np.zeros(t, np.uint64), np.zeros(t, np.uint32), np.zeros(t, np.uint8)
        ng = np.zeros(t, np.uint64) if ngram else None
        if n.value:
            rc = lib().tsm_line_hashes(self._ctx, C.byref(cs), _p(base), _p(lh), _p(le), _p(lf), t, C.byref(n), int(ngram), _p(ng), stream)
            if rc:
                raise TsmError(rc, "tsm_line_hashes")
        m = int(n.value)
        out = (base, lh[:m], le[:m], lf[:m])
        return out + (ng[:m],) if ngram else out

    def reduce(self, flags, repo, case_id, n_repos, n_cases, stream=None):
        flags = np.ascontiguousarray(flags, np.uint8)
        repo = np.ascontiguousarray(repo, np.int32)
        case_id = np.ascontiguousarray(case_id, np.int32)
        n_rows, n_flags = flags.shape
        out = np.zeros((n_flags, n_repos), np.int64)
        cpr = np.zeros(n_repos, np.int64)
        rc = lib().tsm_reduce(self._ctx, _p(flags), _p(repo), _p(case_id), n_rows, n_flags, n_repos, n_cases,
                              _p(out), _p(cpr), stream)
        if rc:
            raise TsmError(rc, "tsm_reduce")
        return out, cpr

    def statements(self, corpus, stream=None):
        """SPEC section 10: (line_base[n+1], line_end[lines], line_kind[lines]); kind 0 blank, 1 statement start, 2 continuation."""
        cs = corpus.c_struct()
        base = np.zeros(corpus.n_files + 1, np.int64)
        n = C.c_int64()
        rc = lib().tsm_statements(self._ctx, C.byref(cs), _p(base), None, None, 0, C.byref(n), stream)
        if rc not in (0, -3):
            raise TsmError(rc, "tsm_statements")
        end = np.zeros(max(n.value, 1), np.uint32)
        kind = np.zeros(max(n.value, 1), np.uint8)
        rc = lib().tsm_statements(self._ctx, C.byref(cs), _p(base), _p(end), _p(kind), n.value, C.byref(n), stream)
        if rc:
            raise TsmError(rc, "tsm_statements")
        return base, end[:n.value], kind[:n.value]

    def diff_pairs(self, olds, news, stream=None, detail=False):
        """S8 churn per pair; with detail=True also the hunks of the canonical edit script (SPEC section 8)."""
        n = olds.n_files
        added = np.zeros(n, np.int64)
        removed = np.zeros(n, np.int64)
        a, b = olds.c_struct(), news.c_struct()
        if not detail:
            rc = lib().tsm_diff_pairs(self._ctx, C.byref(a), C.byref(b), _p(added), _p(removed), stream)
            if rc:
                raise TsmError(rc, "tsm_diff_pairs")
            return added, removed
        det = np.zeros(max(n, 1), DIFF_DETAIL)
        rc = lib().tsm_diff_pairs_detail(self._ctx, C.byref(a), C.byref(b), _p(added), _p(removed), _p(det), stream)
        if rc:
            raise TsmError(rc, "tsm_diff_pairs_detail")
        return added, removed, det[:n]

    def diff_upload(self, olds, news, stream=None):
        """Both sides of the pairs to HBM, kept by the ctx (tsm_diff_upload)."""
        a, b = olds.c_struct(), news.c_struct()
        rc = lib().tsm_diff_upload(self._ctx, C.byref(a), C.byref(b), stream)
        if rc:
            raise TsmError(rc, "tsm_diff_upload")
        self._pairs = olds.n_files
        n = max(self._pairs, 1)                              # results land in pinned memory: D2H at the PCIe rate, no staging copy
        bufs = [host_buffer(n * 8), host_buffer(n * 8), host_buffer(n * DIFF_DETAIL.itemsize)]
        self._diff_pins = [b[1] for b in bufs]
        self._diff_out = (bufs[0][0][:self._pairs * 8].view(np.int64), bufs[1][0][:self._pairs * 8].view(np.int64),
                          bufs[2][0][:n * DIFF_DETAIL.itemsize].view(DIFF_DETAIL))

    def diff_resident(self, detail=True, stream=None):
        """The diff kernels over the resident sides; returns (added, removed[, detail]) - buffers reused across calls."""
        added, removed, det = self._diff_out
        rc = lib().tsm_diff_resident(self._ctx, _p(added), _p(removed), _p(det) if detail else None, stream)
        if rc:
            raise TsmError(rc, "tsm_diff_resident")
        return (added, removed, det[:self._pairs]) if detail else (added, removed)

    def diff_last_ms(self):
        """Device time of the last diff call: [k_scan over both sides, k_diff_small, k_myers + k_myers_trace of the pairs it left over] in ms."""
        ms = (C.c_float * 3)()
        lib().tsm_diff_last_ms(self._ctx, C.byref(ms))
        return [float(x) for x in ms]
