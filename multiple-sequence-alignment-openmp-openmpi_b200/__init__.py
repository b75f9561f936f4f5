"""seqalk-b200: B200-native all-pairs minimum-penalty global alignment.

Host-side mirror of the reference's operator surface for its one hot path
(yangxvlin/multiple-sequence-alignment-openMP-openMPI, ``submit/xuliny-seqalkway.cpp``):

* :func:`getMinimumPenalties`  <- ``getMinimumPenalties(genes, k, pxy, pgap, penalties)`` (:232-364)
* :func:`getMinimumPenalty`    <- ``getMinimumPenalty(x, y, pxy, pgap, xans, yans)`` (:570-663)
* :func:`do_MPI_task`          <- ``do_MPI_task(rank)`` (:369-417) — a no-op here
* :mod:`.driver`               <- the stdin/stdout format of ``main`` (:35-77)
* :mod:`.distributed`          <- the master/worker task farm (:272-362) as one process per GPU

Everything computes through ``libseqalk.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/seqalk.h``).  There is NO CPU fallback: if the library is not built, importing the
compute entry points raises, and every compute call fails without a B200.

The directory name contains hyphens (it is the graft package name); import it with
``importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200")`` or through the
``seqalk_b200`` alias module at the repository root.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

__all__ = [
    "SeqalkError", "lib", "library_path", "getMinimumPenalties", "getMinimumPenalty", "do_MPI_task",
    "get_minimum_penalties", "get_minimum_penalty", "all_pairs_scores", "pair_score", "last_profile", "Job", "JobStats", "partition", "partition_all", "partition_chunks", "chunk_bounds", "task_ij",
    "num_tasks", "fold_chain", "sha512_hex", "init", "shutdown", "device_count", "version",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libseqalk.so"


class SeqalkError(RuntimeError):
    """A libseqalk call failed (message from seqalk_last_error)."""


def library_path() -> str:
    # SEQALK_LIB: another build of the same library (A/B measurements of kernel changes)
    return os.environ.get("SEQALK_LIB") or os.path.join(_HERE, _LIB_NAME)


class _Stats(ctypes.Structure):
    _fields_ = [
        ("cells", ctypes.c_uint64), ("dir_bytes", ctypes.c_uint64), ("batches", ctypes.c_uint32),
        ("fill_launches", ctypes.c_uint32), ("traceback_launches", ctypes.c_uint32),
        ("digest_launches", ctypes.c_uint32), ("fill_ms", ctypes.c_double),
        ("traceback_ms", ctypes.c_double), ("digest_ms", ctypes.c_double), ("wall_ms", ctypes.c_double),
        ("tile_mode", ctypes.c_uint32), ("fill_rows_per_lane", ctypes.c_uint32),
        ("fill_packed16", ctypes.c_uint32), ("self_checked", ctypes.c_uint32),
    ]


# every symbol include/seqalk.h declares: (name, restype, argtypes)
_c = ctypes
_PP = _c.POINTER(_c.c_char_p)
_I32P = _c.POINTER(_c.c_int32)
_I64P = _c.POINTER(_c.c_int64)
_U8P = _c.POINTER(_c.c_uint8)
C_ABI = [
    ("seqalk_init", _c.c_int, [_c.c_int]),
    ("seqalk_shutdown", None, []),
    ("seqalk_device_count", _c.c_int, []),
    ("seqalk_last_error", _c.c_char_p, []),
    ("seqalk_version", _c.c_char_p, []),
    ("seqalk_all_pairs", _c.c_int, [_c.POINTER(_c.c_void_p), _I32P, _c.c_int32, _c.c_int32, _c.c_int32, _I32P, _c.c_char_p]),
    ("seqalk_pair", _c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_int32, _c.c_int32, _c.c_int32, _I32P, _I32P, _I32P]),
    ("seqalk_all_pairs_scores", _c.c_int, [_c.POINTER(_c.c_void_p), _I32P, _c.c_int32, _c.c_int32, _c.c_int32, _I32P]),
    ("seqalk_pair_score", _c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_int32, _c.c_int32, _c.c_int32, _I32P]),
    ("seqalk_profile", _c.c_int, [_c.POINTER(_c.c_double), _c.c_int32]),
    ("seqalk_num_tasks", _c.c_int64, [_c.c_int32]),
    ("seqalk_task_ij", _c.c_int, [_c.c_int64, _I32P, _I32P]),
    ("seqalk_partition", _c.c_int, [_I32P, _c.c_int32, _c.c_int64, _c.c_int64, _c.c_int32, _c.c_int32, _I64P, _I64P]),
    ("seqalk_partition_all", _c.c_int, [_I32P, _c.c_int32, _c.c_int64, _c.c_int64, _c.c_int32, _I64P, _I64P]),
    ("seqalk_chunk_bounds", _c.c_int, [_c.c_int64, _c.c_int32, _I64P]),
    ("seqalk_partition_chunks", _c.c_int, [_I32P, _c.c_int32, _c.c_int32, _c.c_int32, _I64P, _I64P]),
    ("seqalk_job_create", _c.c_int, [_c.c_int32, _c.POINTER(_c.c_void_p), _I32P, _c.c_int32, _c.c_int32, _c.c_int32, _I64P, _c.c_int64, _c.POINTER(_c.c_void_p)]),
    ("seqalk_job_run", _c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p]),
    ("seqalk_job_set_chunks", _c.c_int, [_c.c_void_p, _I64P, _c.c_int32]),
    ("seqalk_job_run_chunk", _c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p, _c.c_void_p]),
    ("seqalk_job_start_async", _c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p]),
    ("seqalk_job_wait_chunk", _c.c_int, [_c.c_void_p, _c.c_int32, _c.c_void_p]),
    ("seqalk_job_finish_async", _c.c_int, [_c.c_void_p]),
    ("seqalk_job_run_scores", _c.c_int, [_c.c_void_p, _c.c_void_p]),
    ("seqalk_job_fetch", _c.c_int, [_c.c_void_p, _I32P, _U8P]),
    ("seqalk_job_stats", _c.c_int, [_c.c_void_p, _c.POINTER(_Stats)]),
    ("seqalk_job_destroy", None, [_c.c_void_p]),
    ("seqalk_fold_chain", _c.c_int, [_U8P, _c.c_int64, _c.c_char_p]),
    ("seqalk_sha512_hex", _c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_char_p]),
    ("seqalk_measure_int_peaks", _c.c_int, [_c.c_int32, _c.POINTER(_c.c_double)]),
]

_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """Loads libseqalk.so (built by ``__graft_entry__.build()`` / ``make -C csrc``).

    Fails loudly when the CUDA library is missing — there is no other implementation."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise SeqalkError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). seqalk-b200 has no CPU fallback.")
        handle = ctypes.CDLL(path)
        for name, restype, argtypes in C_ABI:
            fn = getattr(handle, name)  # AttributeError if the ABI and the header drift apart
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise SeqalkError(lib().seqalk_last_error().decode("utf-8", "replace"))


def _as_bytes(s) -> bytes:
    if isinstance(s, bytes):
        return s
    if isinstance(s, (bytearray, memoryview)):
        return bytes(s)
    if isinstance(s, np.ndarray):
        return s.astype(np.uint8, copy=False).tobytes()
    if isinstance(s, str):
        return s.encode("latin-1")
    raise TypeError(f"sequence must be str/bytes/uint8 array, got {type(s)!r}")


class _SeqArgs:
    """Keeps the byte buffers alive and exposes (char**, int32*) for the C ABI."""

    def __init__(self, genes: Sequence):
        self.bufs: List[bytes] = [_as_bytes(g) for g in genes]
        self.k = len(self.bufs)
        self.lens = np.array([len(b) for b in self.bufs], dtype=np.int32)
        self.ptrs = (ctypes.c_void_p * max(self.k, 1))()
        # pointers into the bytes objects themselves (kept alive by self.bufs): no copy — copying
        # the 10 MB of BASELINE config 4 into 1000 ctypes buffers cost 14 ms per call
        self._empty = ctypes.create_string_buffer(1)
        for i, b in enumerate(self.bufs):
            self.ptrs[i] = (ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p).value if len(b)
                            else ctypes.addressof(self._empty))

    @property
    def lens_p(self):
        return self.lens.ctypes.data_as(_I32P)


# ------------------------------------------------------------------ runtime

def init(ngpus: int = 0) -> None:
    _check(lib().seqalk_init(int(ngpus)))


def shutdown() -> None:
    lib().seqalk_shutdown()


def device_count() -> int:
    return int(lib().seqalk_device_count())


def version() -> str:
    return lib().seqalk_version().decode()


# ------------------------------------------------------------------ reference operator surface

def get_minimum_penalties(genes: Sequence, pxy: int, pgap: int) -> Tuple[str, np.ndarray]:
    """All k(k-1)/2 pairs in the reference's task order.

    Returns ``(digest, penalties)``: the chained SHA-512 digest (128 hex chars) and the int32
    penalties ``penalties[id]`` for ``id`` enumerating ``for i in 1..k-1: for j in 0..i-1``."""
    a = _SeqArgs(genes)
    T = a.k * (a.k - 1) // 2
    penalties = np.zeros(max(T, 1), dtype=np.int32)
    digest = ctypes.create_string_buffer(129)
    _check(lib().seqalk_all_pairs(a.ptrs, a.lens_p, a.k, int(pxy), int(pgap),
                                  penalties.ctypes.data_as(_I32P), digest))
    return digest.value.decode("ascii"), penalties[:T]


def all_pairs_scores(genes: Sequence, pxy: int, pgap: int) -> np.ndarray:
    """Score-only form: the penalties ``dp[m][n]`` of all pairs (task order) from the DP fill,
    without alignment strings or digest (``seqalk_all_pairs_scores``)."""
    a = _SeqArgs(genes)
    T = a.k * (a.k - 1) // 2
    penalties = np.zeros(max(T, 1), dtype=np.int32)
    _check(lib().seqalk_all_pairs_scores(a.ptrs, a.lens_p, a.k, int(pxy), int(pgap),
                                         penalties.ctypes.data_as(_I32P)))
    return penalties[:T]


def pair_score(x, y, pxy: int, pgap: int) -> int:
    """Score-only form of one pair (``seqalk_pair_score``): ``dp[m][n]``."""
    xb, yb = _as_bytes(x), _as_bytes(y)
    pen = ctypes.c_int32(0)
    xbuf = ctypes.create_string_buffer(xb, max(len(xb), 1))
    ybuf = ctypes.create_string_buffer(yb, max(len(yb), 1))
    _check(lib().seqalk_pair_score(ctypes.cast(xbuf, ctypes.c_void_p), len(xb), ctypes.cast(ybuf, ctypes.c_void_p),
                                   len(yb), int(pxy), int(pgap), ctypes.byref(pen)))
    return int(pen.value)


PROFILE_NAMES = ["call_ms", "setup_ms", "gpu_span_ms", "fold_exposed_ms", "gather_d2h_ms", "kernel_ms_max_gpu",
                 "kernel_ms_mean_gpu", "gpus", "chunks", "fill_ms_sum", "traceback_ms_sum", "digest_ms_sum"]


def last_profile() -> dict:
    """Phase times of the last get_minimum_penalties call of this process (``seqalk_profile``)."""
    out = (ctypes.c_double * len(PROFILE_NAMES))()
    _check(lib().seqalk_profile(out, len(PROFILE_NAMES)))
    return dict(zip(PROFILE_NAMES, [float(v) for v in out]))


def getMinimumPenalties(genes: Sequence, k: int, pxy: int, pgap: int, penalties) -> str:
    """Reference signature (submit/xuliny-seqalkway.cpp:13): fills ``penalties`` in place
    (any writable int sequence of length >= k(k-1)/2) and returns the digest string."""
    if k != len(genes):
        genes = list(genes)[:k]
    digest, pen = get_minimum_penalties(genes, pxy, pgap)
    for t in range(len(pen)):
        penalties[t] = int(pen[t])
    return digest


def get_minimum_penalty(x, y, pxy: int, pgap: int) -> Tuple[int, np.ndarray, np.ndarray]:
    """One pair; returns ``(penalty, xans, yans)`` with the reference's right-aligned int arrays
    of length m+n+1 (index 0 unused, '_' = 95 pad in front)."""
    xb, yb = _as_bytes(x), _as_bytes(y)
    m, n = len(xb), len(yb)
    xans = np.zeros(m + n + 1, dtype=np.int32)
    yans = np.zeros(m + n + 1, dtype=np.int32)
    pen = ctypes.c_int32(0)
    xbuf = ctypes.create_string_buffer(xb, max(m, 1))
    ybuf = ctypes.create_string_buffer(yb, max(n, 1))
    _check(lib().seqalk_pair(ctypes.cast(xbuf, ctypes.c_void_p), m, ctypes.cast(ybuf, ctypes.c_void_p), n,
                             int(pxy), int(pgap), xans.ctypes.data_as(_I32P), yans.ctypes.data_as(_I32P),
                             ctypes.byref(pen)))
    return int(pen.value), xans, yans


def getMinimumPenalty(x, y, pxy: int, pgap: int, xans, yans) -> int:
    """Reference signature (submit/xuliny-seqalkway.cpp:14): fills ``xans``/``yans`` (index 1..m+n)."""
    pen, xa, ya = get_minimum_penalty(x, y, pxy, pgap)
    for a in range(1, len(xa)):
        xans[a] = int(xa[a])
        yans[a] = int(ya[a])
    return pen


def do_MPI_task(rank: int) -> None:
    """Reference signature (submit/xuliny-seqalkway.cpp:15).  Non-root ranks have nothing to
    do: the GPUs are driven from the calling process (or one process per GPU, see
    :mod:`.distributed`)."""
    return None


def alignment_strings(xans: np.ndarray, yans: np.ndarray) -> Tuple[bytes, bytes]:
    """The reference's strip scan (do_task, :191-214): drop the common leading '_' pad."""
    l = len(xans) - 1
    start = 1
    for a in range(l, 0, -1):
        if (xans[a] & 0xFF) == 95 and (yans[a] & 0xFF) == 95:
            start = a + 1
            break
    return (bytes(int(v) & 0xFF for v in xans[start:]), bytes(int(v) & 0xFF for v in yans[start:]))


# ------------------------------------------------------------------ sharded execution

def num_tasks(k: int) -> int:
    return int(lib().seqalk_num_tasks(int(k)))


def task_ij(task_id: int) -> Tuple[int, int]:
    i, j = ctypes.c_int32(), ctypes.c_int32()
    _check(lib().seqalk_task_ij(int(task_id), ctypes.byref(i), ctypes.byref(j)))
    return int(i.value), int(j.value)


def partition(lens: Iterable[int], nranks: int, rank: int, id_begin: int = 0,
              id_end: Optional[int] = None) -> np.ndarray:
    """Cell-balanced (LPT) shard of task ids [id_begin, id_end) for ``rank``; increasing ids."""
    lens_a = np.ascontiguousarray(np.asarray(list(lens), dtype=np.int32))
    k = len(lens_a)
    if id_end is None:
        id_end = k * (k - 1) // 2
    out = np.zeros(max(id_end - id_begin, 1), dtype=np.int64)
    cnt = ctypes.c_int64(0)
    _check(lib().seqalk_partition(lens_a.ctypes.data_as(_I32P), k, int(id_begin), int(id_end), int(nranks),
                                  int(rank), out.ctypes.data_as(_I64P), ctypes.byref(cnt)))
    return out[: cnt.value].copy()


def partition_all(lens: Iterable[int], nranks: int, id_begin: int = 0,
                  id_end: Optional[int] = None) -> List[np.ndarray]:
    """The shards of every rank (list of increasing id arrays), from one LPT pass."""
    lens_a = np.ascontiguousarray(np.asarray(list(lens), dtype=np.int32))
    k = len(lens_a)
    if id_end is None:
        id_end = k * (k - 1) // 2
    out = np.zeros(max(id_end - id_begin, 1), dtype=np.int64)
    counts = np.zeros(nranks, dtype=np.int64)
    _check(lib().seqalk_partition_all(lens_a.ctypes.data_as(_I32P), k, int(id_begin), int(id_end), int(nranks),
                                      out.ctypes.data_as(_I64P), counts.ctypes.data_as(_I64P)))
    ends = np.cumsum(counts)
    return [out[e - c:e].copy() for c, e in zip(counts, ends)]


def chunk_bounds(ntasks: int, nchunks: int) -> List[int]:
    """Boundaries of the id chunks (tapered tail, see include/seqalk.h)."""
    out = np.zeros(nchunks + 1, dtype=np.int64)
    _check(lib().seqalk_chunk_bounds(int(ntasks), int(nchunks), out.ctypes.data_as(_I64P)))
    return [int(v) for v in out]


def partition_chunks(lens: Iterable[int], nranks: int, nchunks: int) -> List[List[np.ndarray]]:
    """LPT shards of ``nchunks`` consecutive id ranges for every rank: result[c][r] (ids increasing)."""
    lens_a = np.ascontiguousarray(np.asarray(list(lens), dtype=np.int32))
    k = len(lens_a)
    T = k * (k - 1) // 2
    out = np.zeros(max(T, 1), dtype=np.int64)
    counts = np.zeros(nchunks * nranks, dtype=np.int64)
    _check(lib().seqalk_partition_chunks(lens_a.ctypes.data_as(_I32P), k, int(nranks), int(nchunks),
                                         out.ctypes.data_as(_I64P), counts.ctypes.data_as(_I64P)))
    ends = np.cumsum(counts)
    flat = [out[e - c:e] for c, e in zip(counts, ends)]
    return [flat[c * nranks:(c + 1) * nranks] for c in range(nchunks)]


class JobStats(dict):
    __getattr__ = dict.__getitem__


class Job:
    """Tasks resident on one GPU (sequences uploaded, batches planned).

    ``run()`` executes fill -> traceback -> digest for every task with inputs already in HBM;
    results stay on the device (job-owned arrays, or caller-provided device pointers such as
    ``torch.Tensor.data_ptr()``) until :meth:`fetch`."""

    def __init__(self, genes: Sequence, pxy: int, pgap: int, task_ids: Optional[np.ndarray] = None,
                 device: int = 0):
        self._args = _SeqArgs(genes)
        self._h = ctypes.c_void_p()
        if task_ids is not None:
            ids = np.ascontiguousarray(np.asarray(task_ids, dtype=np.int64))
            idp, n = ids.ctypes.data_as(_I64P), len(ids)
            self.ntasks = n
        else:
            ids, idp, n = None, None, 0
            self.ntasks = self._args.k * (self._args.k - 1) // 2
        _check(lib().seqalk_job_create(int(device), self._args.ptrs, self._args.lens_p, self._args.k, int(pxy),
                                       int(pgap), idp, n, ctypes.byref(self._h)))

    def run(self, penalties_dev: int = 0, digests_dev: int = 0) -> None:
        _check(lib().seqalk_job_run(self._h, ctypes.c_void_p(penalties_dev or None),
                                    ctypes.c_void_p(digests_dev or None)))

    def set_chunks(self, chunk_ends: Sequence[int]) -> None:
        """Cut the job into chunks of consecutive result slots (cumulative end indices)."""
        ends = np.ascontiguousarray(np.asarray(chunk_ends, dtype=np.int64))
        _check(lib().seqalk_job_set_chunks(self._h, ends.ctypes.data_as(_I64P), len(ends)))

    def run_chunk(self, chunk: int, penalties_dev: int = 0, digests_dev: int = 0) -> None:
        """run() for one chunk; the device pointers address the whole job's result arrays."""
        _check(lib().seqalk_job_run_chunk(self._h, int(chunk), ctypes.c_void_p(penalties_dev or None),
                                          ctypes.c_void_p(digests_dev or None)))

    def start_async(self, penalties_dev: int = 0, digests_dev: int = 0) -> None:
        """All chunks back to back, enqueued by a library thread (see wait_chunk)."""
        _check(lib().seqalk_job_start_async(self._h, ctypes.c_void_p(penalties_dev or None),
                                            ctypes.c_void_p(digests_dev or None)))

    def wait_chunk(self, chunk: int, consumer_stream: int = 0) -> None:
        """Order ``consumer_stream`` (raw cudaStream_t, 0 = block the host) after chunk's results."""
        _check(lib().seqalk_job_wait_chunk(self._h, int(chunk), ctypes.c_void_p(consumer_stream or None)))

    def finish_async(self) -> None:
        _check(lib().seqalk_job_finish_async(self._h))

    def run_scores(self, penalties_dev: int = 0) -> None:
        """Score-only pass (penalties from the fill, no traceback / digest)."""
        _check(lib().seqalk_job_run_scores(self._h, ctypes.c_void_p(penalties_dev or None)))

    def fetch_penalties(self) -> np.ndarray:
        pen = np.zeros(max(self.ntasks, 1), dtype=np.int32)
        _check(lib().seqalk_job_fetch(self._h, pen.ctypes.data_as(_I32P), None))
        return pen[: self.ntasks]

    def fetch(self) -> Tuple[np.ndarray, np.ndarray]:
        pen = np.zeros(max(self.ntasks, 1), dtype=np.int32)
        dig = np.zeros((max(self.ntasks, 1), 64), dtype=np.uint8)
        _check(lib().seqalk_job_fetch(self._h, pen.ctypes.data_as(_I32P), dig.ctypes.data_as(_U8P)))
        return pen[: self.ntasks], dig[: self.ntasks]

    def stats(self) -> JobStats:
        s = _Stats()
        _check(lib().seqalk_job_stats(self._h, ctypes.byref(s)))
        return JobStats({name: getattr(s, name) for name, _ in _Stats._fields_})

    def close(self) -> None:
        if self._h:
            lib().seqalk_job_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fold_chain(pair_digests: np.ndarray, state: str = "") -> str:
    """Folds per-task digests (n x 64 uint8, task order) into the chain state."""
    d = np.ascontiguousarray(pair_digests, dtype=np.uint8).reshape(-1, 64)
    buf = ctypes.create_string_buffer(state.encode("ascii"), 129)
    _check(lib().seqalk_fold_chain(d.ctypes.data_as(_U8P), len(d), buf))
    return buf.value.decode("ascii")


INT_PEAK_NAMES = ["lop3", "vimnmx3", "imad", "lop3+imad", "viaddmnmx", "shf", "lop3+vimnmx3", "fill_mix",
                  "idp4a", "idp4a+vimnmx3", "idp4a+imad", "prmt", "vimnmx3.u16x2", "prmt+imad+vimnmx3.u16x2",
                  "idp4a+vimnmx3 distinct operands", "reserved"]


def measure_int_peaks(device: int = 0) -> dict:
    """Issue-rate microbenchmarks (lane-ops/s) — the integer roofline denominators."""
    r = (ctypes.c_double * 16)()
    if lib().seqalk_measure_int_peaks(int(device), r) != 0:
        raise SeqalkError("seqalk_measure_int_peaks failed (no usable GPU?)")
    return dict(zip(INT_PEAK_NAMES, [float(v) for v in r]))


def sha512_hex(data: bytes) -> str:
    out = ctypes.create_string_buffer(129)
    b = ctypes.create_string_buffer(data, max(len(data), 1))
    _check(lib().seqalk_sha512_hex(ctypes.cast(b, ctypes.c_void_p), len(data), out))
    return out.value.decode("ascii")
