#!/usr/bin/env python
"""Generates tests/golden/fullsize.json: CPU-oracle goldens for BASELINE configs 3, 4, 5 at
FULL size (the -m gpu tests assert them; nothing here needs a GPU or /root/reference).

    python tests/golden/make_golden_fullsize.py [cfg5] [cfg4] [cfg3]    (default: all three)

  cfg5  the single 200 000 x 200 000 pair: penalty + problemhash + final digest (oracle.pair,
        one core, ~10 GB of packed directions, minutes).
  cfg4  k=1000: a seeded, stratified sample of >= 2 000 of the 499 500 tasks — strata = deciles
        of the cell count, plus the 64 most skewed length ratios, the 32 largest and the 32
        smallest tasks: per task penalty + problemhash.
  cfg3  k=200 x 10 000: the whole job — digest, all 19 900 penalties, every 50th problemhash and
        a SHA-512 over all problemhashes (~1.99e12 cells: about 30-40 min on 8 cores).

The oracle follows testing12/seqalign-mpi-skeleton.cpp:117-175 (order, orientation, chain) and
:186-280 (fill, traceback, tails) and is itself pinned by the reference's captured outputs
(tests/golden/make_golden.py).  Entries already present in fullsize.json are kept unless named
on the command line.
"""
import hashlib
import importlib
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import binding as oracle  # noqa: E402

workloads = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200.workloads")
OUT = os.path.join(HERE, "fullsize.json")


def task_ij(t):
    i = int((1 + (1 + 8 * t) ** 0.5) / 2)
    while i * (i - 1) // 2 > t:
        i -= 1
    while (i + 1) * i // 2 <= t:
        i += 1
    return i, t - i * (i - 1) // 2


def cfg4_sample_ids(lens, per_stratum=190, seed=4):
    """Deterministic stratified sample of task ids (documented in the module docstring)."""
    k = len(lens)
    L = np.asarray(lens, dtype=np.int64)
    ii, jj = np.tril_indices(k, -1)  # row-major: i = 1.., j < i  == the reference's task order
    cells = L[ii] * L[jj]
    ratio = np.maximum(L[ii], L[jj]) / np.minimum(L[ii], L[jj])
    order = np.argsort(cells, kind="stable")
    rng = np.random.RandomState(seed)
    picked = set()
    T = len(cells)
    for d in range(10):
        stratum = order[T * d // 10: T * (d + 1) // 10]
        picked.update(int(v) for v in rng.choice(stratum, size=per_stratum, replace=False))
    picked.update(int(v) for v in np.argsort(-ratio, kind="stable")[:64])
    picked.update(int(v) for v in order[-32:])
    picked.update(int(v) for v in order[:32])
    return sorted(picked)


def main():
    which = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg5", "cfg4", "cfg3"]
    oracle.build(with_ref=False)
    doc = json.load(open(OUT)) if os.path.exists(OUT) else {}

    def save():
        json.dump(doc, open(OUT, "w"), indent=None, separators=(",", ":"))

    if "cfg5" in which:
        pxy, pgap, genes = workloads.make("cfg5")
        t0 = time.time()
        pen, a1, a2, ph = oracle.pair(genes[1], genes[0], pxy, pgap)  # task (i=1, j=0): x = genes[1]
        doc["cfg5"] = {"pxy": pxy, "pgap": pgap, "k": 2, "lens": [len(g) for g in genes], "penalty": pen,
                       "alignment_length": len(a1), "problemhash": ph,
                       "digest": hashlib.sha512(ph.encode()).hexdigest(),
                       "align1_sha512": hashlib.sha512(a1).hexdigest(),
                       "align2_sha512": hashlib.sha512(a2).hexdigest(),
                       "source": "oracle.pair, %.0f s" % (time.time() - t0)}
        save()
        print("cfg5", pen, ph[:16], doc["cfg5"]["source"], flush=True)

    if "cfg4" in which:
        pxy, pgap, genes = workloads.make("cfg4")
        ids = cfg4_sample_ids([len(g) for g in genes])
        t0 = time.time()

        def one(t):
            i, j = task_ij(t)
            pen, _, _, ph = oracle.pair(genes[i], genes[j], pxy, pgap)
            return pen, ph

        with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:  # ctypes releases the GIL
            res = list(ex.map(one, ids))
        doc["cfg4"] = {"pxy": pxy, "pgap": pgap, "k": len(genes), "cells": workloads.total_cells(genes),
                       "task_ids": ids, "penalties": [r[0] for r in res], "problemhashes": [r[1] for r in res],
                       "sample": "10 cell-count deciles x 190 + 64 most skewed + 32 largest + 32 smallest (seed 4)",
                       "source": "oracle.pair per task, %.0f s" % (time.time() - t0)}
        save()
        print("cfg4", len(ids), "tasks", doc["cfg4"]["source"], flush=True)

    if "cfg3" in which:
        pxy, pgap, genes = workloads.make("cfg3")
        t0 = time.time()
        digest, pen, hashes = oracle.all_pairs(genes, pxy, pgap, want_pair_hashes=True)
        doc["cfg3"] = {"pxy": pxy, "pgap": pgap, "k": len(genes), "cells": workloads.total_cells(genes),
                       "digest": digest, "penalties": [int(v) for v in pen],
                       "problemhash_every_50": hashes[::50],
                       "problemhashes_sha512": hashlib.sha512("".join(hashes).encode()).hexdigest(),
                       "source": "oracle.all_pairs, %.0f s" % (time.time() - t0)}
        save()
        print("cfg3", digest[:16], doc["cfg3"]["source"], flush=True)


if __name__ == "__main__":
    main()
