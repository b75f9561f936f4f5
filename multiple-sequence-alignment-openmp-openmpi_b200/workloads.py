"""Synthetic workloads of BASELINE.json / SURVEY.md section 8d and the reference's text format.

Generator: SplitMix64, ``base = "ACGT"[next() >> 62]``, sequences generated in index order.
The stream is position-addressable (state_n = seed + n*gamma), so generation is vectorised.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def splitmix64(seed: int, start: int, count: int) -> np.ndarray:
    """Outputs number start+1 .. start+count of the SplitMix64 stream seeded with ``seed``
    (the first next() is output number 1)."""
    with np.errstate(over="ignore"):
        n = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + n * _GAMMA
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _bases(seed: int, start: int, lens: Sequence[int]) -> List[bytes]:
    total = int(sum(lens))
    codes = (splitmix64(seed, start, total) >> np.uint64(62)).astype(np.intp)
    flat = _ACGT[codes]
    out, pos = [], 0
    for L in lens:
        out.append(flat[pos:pos + L].tobytes())
        pos += L
    return out


def uniform(k: int, length: int, seed: int) -> List[bytes]:
    """k sequences of equal length, bases i.i.d. uniform over ACGT (configs 3 and 5)."""
    return _bases(seed, 0, [length] * k)


def mixed(k: int, seed: int, lo: int = 1000, hi: int = 20000) -> List[bytes]:
    """Config 4: len_i = lo + next() % (hi-lo+1) drawn first for all i, then the bases."""
    lens = (lo + (splitmix64(seed, 0, k) % np.uint64(hi - lo + 1))).astype(np.int64)
    return _bases(seed, k, [int(v) for v in lens])


def _golden(name: str):
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return parse_input(open(os.path.join(root, "tests", "golden", name), "rb").read())


# name -> (description, factory)
CONFIGS = {
    "cfg1": ("mseq1.dat (reference fixture, k=9)", lambda: _golden("mseq1.dat")),
    "cfg2": ("mseq-big13-example.txt (reference fixture, k=13, 30k-90k, 278.5 G cells)",
             lambda: _golden("mseq-big13-example.txt")),
    "cfg3": ("synthetic k=200, length 10k uniform ACGT, all-pairs (seed 200)", lambda: (3, 2, uniform(200, 10000, 200))),
    "cfg4": ("synthetic k=1000, lengths 1k-20k mixed, all-pairs (seed 1000)", lambda: (3, 2, mixed(1000, 1000))),
    "cfg5": ("synthetic k=2, length 200k single pair (seed 2)", lambda: (3, 2, uniform(2, 200000, 2))),
    # reduced shapes of the same generators for quick checks (not BASELINE configs)
    "cfg3-small": ("synthetic k=24, length 10k uniform ACGT (seed 200)", lambda: (3, 2, uniform(24, 10000, 200))),
    "cfg4-small": ("synthetic k=100, lengths 1k-20k mixed (seed 1000)", lambda: (3, 2, mixed(100, 1000))),
}


def make(name: str) -> Tuple[int, int, List[bytes]]:
    """Returns (pxy, pgap, genes) of a named workload."""
    return CONFIGS[name][1]()


def total_cells(genes: Sequence[bytes]) -> int:
    """Sum over pairs i>j of len_i*len_j (Python ints: no overflow)."""
    lens = [len(g) for g in genes]
    s, acc = 0, 0
    for L in lens:
        acc += L * s
        s += L
    return acc


def parse_input(text: bytes) -> Tuple[int, int, List[bytes]]:
    """stdin format of the reference's main (submit/xuliny-seqalkway.cpp:44-48):
    ``pxy pgap k`` then k whitespace-delimited tokens."""
    tok = text.split()
    pxy, pgap, k = int(tok[0]), int(tok[1]), int(tok[2])
    genes = tok[3:3 + k]
    if len(genes) != k:
        raise ValueError(f"expected {k} sequences, found {len(genes)}")
    return pxy, pgap, genes


def format_input(pxy: int, pgap: int, genes: Sequence[bytes]) -> bytes:
    return b"\n".join([str(pxy).encode(), str(pgap).encode(), str(len(genes)).encode(), *genes]) + b"\n"


def format_output(time_us: int, digest: str, penalties: Sequence[int]) -> str:
    """stdout format of the reference's main (:62-70)."""
    return f"Time: {time_us} us\n{digest}\n" + "".join(f"{int(p)} " for p in penalties) + "\n"
