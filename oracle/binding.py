"""ctypes loader for the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing in the product package imports it."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle.so")
REF_SEQ = os.path.join(_HERE, "_ref", "ref_seq")   # unmodified sequential skeleton + mpi shim
REF_PAR = os.path.join(_HERE, "_ref", "ref_par")   # unmodified submission + mpi shim ("np=1")
_lib = None


def build(with_ref: bool = True) -> None:
    """make the oracle library; and the reference binaries when /root/reference exists."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)
    if with_ref and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build(with_ref=False)
        h = ctypes.CDLL(LIB)
        h.oracle_pair.restype = ctypes.c_int
        h.oracle_pair.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_char_p,
                                  ctypes.POINTER(ctypes.c_int), ctypes.c_char_p]
        h.oracle_pair_score.restype = ctypes.c_int
        h.oracle_pair_score.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int]
        h.oracle_all_pairs.restype = ctypes.c_int
        h.oracle_all_pairs.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p,
                                       ctypes.c_char_p, ctypes.c_int]
        h.oracle_sha512_hex.restype = None
        h.oracle_sha512_hex.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        _lib = h
    return _lib


def pair(x: bytes, y: bytes, pxy: int, pgap: int) -> Tuple[int, bytes, bytes, str]:
    """(penalty, align1, align2, problemhash) of one pair (x rows, y columns)."""
    m, n = len(x), len(y)
    pen, alen = ctypes.c_int(0), ctypes.c_int(0)
    a1 = ctypes.create_string_buffer(m + n + 1)
    a2 = ctypes.create_string_buffer(m + n + 1)
    hh = ctypes.create_string_buffer(129)
    rc = lib().oracle_pair(x, m, y, n, pxy, pgap, ctypes.byref(pen), a1, a2, ctypes.byref(alen), hh)
    if rc:
        raise RuntimeError("oracle_pair failed")
    return pen.value, a1.raw[:alen.value], a2.raw[:alen.value], hh.value.decode()


def pair_score(x: bytes, y: bytes, pxy: int, pgap: int) -> int:
    return int(lib().oracle_pair_score(x, len(x), y, len(y), pxy, pgap))


def all_pairs(genes: Sequence[bytes], pxy: int, pgap: int, nthreads: int = 0,
              want_pair_hashes: bool = False) -> Tuple[str, np.ndarray, Optional[List[str]]]:
    """(digest, penalties, per-pair problemhash list or None) in the reference's task order."""
    k = len(genes)
    T = k * (k - 1) // 2
    arr = (ctypes.c_char_p * k)(*genes)
    lens = (ctypes.c_int * k)(*[len(g) for g in genes])
    pen = (ctypes.c_int * max(T, 1))()
    dig = ctypes.create_string_buffer(129)
    ph = ctypes.create_string_buffer(129 * max(T, 1)) if want_pair_hashes else None
    rc = lib().oracle_all_pairs(arr, lens, k, pxy, pgap, pen, dig, ph, nthreads)
    if rc:
        raise RuntimeError("oracle_all_pairs failed")
    hashes = None
    if ph is not None:
        raw = ph.raw
        hashes = [raw[129 * t:129 * t + 128].decode() for t in range(T)]
    return dig.value.decode(), np.array(pen[:T], dtype=np.int32), hashes


def sha512_hex(data: bytes) -> str:
    out = ctypes.create_string_buffer(129)
    lib().oracle_sha512_hex(data, len(data), out)
    return out.value.decode()


def run_reference(binary: str, stdin_bytes: bytes, timeout: Optional[float] = None) -> Tuple[int, str, List[int]]:
    """Runs oracle/_ref/ref_seq or ref_par (the unmodified reference) on a stdin document;
    returns (time_us, digest, penalties)."""
    env = dict(os.environ, OMP_WAIT_POLICY="passive")
    # Packet answers[total] is a stack VLA in the submission (submit/xuliny-seqalkway.cpp:305)
    cmd = f"ulimit -s unlimited 2>/dev/null; exec {binary}"
    p = subprocess.run(["bash", "-c", cmd], input=stdin_bytes, stdout=subprocess.PIPE, env=env, check=True,
                       timeout=timeout)
    lines = p.stdout.decode().splitlines()
    t_us = int(lines[0].split()[1])
    return t_us, lines[1].strip(), [int(v) for v in lines[2].split()]
