"""CPU: the C-ABI library loads, exports what include/seqalk.h declares, and its host-side
logic (task order, LPT shard, SHA-512, chain fold, workload generators, text format) is right.
No compute call is made here (there is no GPU)."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(seqalk):
    header = open(os.path.join(ROOT, "include", "seqalk.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(seqalk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 17
    handle = ctypes.CDLL(seqalk.library_path())
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in seqalk.h but not exported"
    assert declared == {n for n, _, _ in seqalk.C_ABI}


def test_no_cpu_fallback(seqalk):
    if _has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(seqalk.SeqalkError, match="no CPU fallback"):
        seqalk.get_minimum_penalties([b"ACGT", b"AGGT"], 3, 2)
    with pytest.raises(seqalk.SeqalkError, match="no CPU fallback"):
        seqalk.Job([b"ACGT", b"AGGT"], 3, 2)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "multiple-sequence-alignment-openmp-openmpi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
    for f in ["seqalk.h", "seqalk.hpp"]:
        assert "oracle" not in open(os.path.join(ROOT, "include", f)).read().lower()


def test_task_order_matches_reference(seqalk):
    # for i in 1..k-1: for j in 0..i-1  (submit/xuliny-seqalkway.cpp:282-287)
    tid = 0
    for i in range(1, 40):
        for j in range(i):
            assert seqalk.task_ij(tid) == (i, j)
            tid += 1
    assert seqalk.num_tasks(40) == tid
    assert seqalk.num_tasks(1000) == 499500
    assert seqalk.task_ij(499499) == (999, 998)


def test_partition_is_a_balanced_cover(seqalk, workloads):
    pxy, pgap, genes = workloads.make("cfg4")
    lens = np.array([len(g) for g in genes], dtype=np.int64)
    T = seqalk.num_tasks(len(genes))
    ii = np.repeat(np.arange(1000), np.arange(1000))
    jj = np.concatenate([np.arange(i) for i in range(1000)])
    cells = lens[ii] * lens[jj]
    assert int(cells.sum()) == workloads.total_cells(genes) == 54521576330990
    for nranks in (2, 8):
        parts = seqalk.partition_all(lens, nranks)
        allids = np.concatenate(parts)
        assert len(allids) == T and np.array_equal(np.sort(allids), np.arange(T))
        loads = np.array([cells[p].sum() for p in parts], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.0001
        for r, p in enumerate(parts):
            assert np.all(np.diff(p) > 0)
            assert np.array_equal(p, seqalk.partition(lens, nranks, r))
    # sub-range
    parts = seqalk.partition_all(lens, 4, 1000, 5000)
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(1000, 5000))


def test_host_sha512_and_chain(seqalk):
    for n in [0, 1, 111, 112, 127, 128, 129, 255, 256, 257, 5000]:
        data = bytes((7 * i + n) & 0xFF for i in range(n))
        assert seqalk.sha512_hex(data) == hashlib.sha512(data).hexdigest()
    rng = np.random.default_rng(3)
    digs = rng.integers(0, 256, size=(50, 64), dtype=np.uint8)
    h = ""
    for d in digs:  # alignmentHash = H(alignmentHash ++ problemhash)  (:334-337)
        h = hashlib.sha512((h + d.tobytes().hex()).encode()).hexdigest()
    assert seqalk.fold_chain(digs) == h
    # folding in pieces carries the state
    s = seqalk.fold_chain(digs[:17])
    s = seqalk.fold_chain(digs[17:17], s)
    s = seqalk.fold_chain(digs[17:], s)
    assert s == h
    assert seqalk.fold_chain(digs[:0]) == ""


def test_splitmix_generators(workloads):
    def scalar(seed, count):
        s, out, M = seed, [], (1 << 64) - 1
        for _ in range(count):
            s = (s + 0x9E3779B97F4A7C15) & M
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
            out.append(z ^ (z >> 31))
        return out

    ref = scalar(200, 64)
    assert [int(v) for v in workloads.splitmix64(200, 0, 64)] == ref
    assert [int(v) for v in workloads.splitmix64(200, 10, 20)] == ref[10:30]
    g = workloads.uniform(3, 20, 200)
    stream = scalar(200, 60)
    assert g[1] == bytes(b"ACGT"[v >> 62] for v in stream[20:40])
    genes = workloads.mixed(5, 1000)
    st = scalar(1000, 5 + sum(len(x) for x in genes))
    assert [len(x) for x in genes] == [1000 + v % 19001 for v in st[:5]]
    assert genes[0] == bytes(b"ACGT"[v >> 62] for v in st[5:5 + len(genes[0])])


def test_text_format_roundtrip(workloads):
    doc = b"3\n2\n3\nAGGGCT\nAGGCA   AAAGGGCT\n"
    pxy, pgap, genes = workloads.parse_input(doc)
    assert (pxy, pgap, genes) == (3, 2, [b"AGGGCT", b"AGGCA", b"AAAGGGCT"])
    assert workloads.parse_input(workloads.format_input(pxy, pgap, genes)) == (pxy, pgap, genes)
    assert workloads.format_output(12, "ab", [5, 4, 9]) == "Time: 12 us\nab\n5 4 9 \n"


def test_alignment_strings_strip_scan(seqalk):
    x = np.array([0, 95, 95, 65, 95], dtype=np.int32)
    y = np.array([0, 95, 95, 95, 67], dtype=np.int32)
    assert seqalk.alignment_strings(x, y) == (b"A_", b"_C")


def test_partition_chunks_matches_per_chunk_partition(seqalk):
    lens = [100, 2000, 350, 900, 1200, 60, 777, 5000, 41, 3000]
    T = seqalk.num_tasks(len(lens))
    for nranks, nchunks in [(1, 1), (3, 4), (8, 2)]:
        chunks = seqalk.partition_chunks(lens, nranks, nchunks)
        assert len(chunks) == nchunks and all(len(c) == nranks for c in chunks)
        seen = np.concatenate([a for c in chunks for a in c])
        assert np.array_equal(np.sort(seen), np.arange(T))
        b = seqalk.chunk_bounds(T, nchunks)
        assert b[0] == 0 and b[-1] == T and all(x <= y for x, y in zip(b, b[1:]))
        for c in range(nchunks):
            ref = seqalk.partition_all(lens, nranks, b[c], b[c + 1])
            assert all(np.array_equal(a, b) for a, b in zip(ref, chunks[c]))


def test_chunk_bounds_shrink_geometrically(seqalk):
    """Chunk sizes shrink by 0.8 per chunk (csrc/api.cu chunk_bounds): the results in flight — and the
    chain fold left when the GPUs finish — shrink with them; small jobs get equal chunks."""
    T = 499500
    for n in (4, 10, 14):
        b = seqalk.chunk_bounds(T, n)
        sizes = np.diff(np.asarray(b)).astype(float)
        assert b[0] == 0 and b[-1] == T and (sizes > 0).all()
        assert np.allclose(sizes[1:] / sizes[:-1], 0.8, atol=0.01)
    assert sum(np.diff(seqalk.chunk_bounds(T, 14))[-3:]) < 0.05 * T
    assert seqalk.chunk_bounds(100, 4) == [0, 25, 50, 75, 100]


def test_every_knob_in_the_sources_is_documented():
    """INTEGRATION.md section 4 lists every SEQALK_* environment knob the library or the package reads."""
    import glob
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg_dir = os.path.join(root, "multiple-sequence-alignment-openmp-openmpi_b200")
    knobs = set()
    for path in glob.glob(os.path.join(pkg_dir, "csrc", "*")) + glob.glob(os.path.join(pkg_dir, "*.py")):
        if os.path.isfile(path) and path.endswith((".cu", ".cpp", ".h", ".cuh", ".py")):
            text = open(path, errors="replace").read()
            knobs.update(re.findall(r'getenv\("(SEQALK_[A-Z0-9_]+)"\)', text))
            knobs.update(re.findall(r'environ(?:\.get\(|\[)"(SEQALK_[A-Z0-9_]+)"', text))
    assert len(knobs) > 20
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    missing = sorted(k for k in knobs if k not in doc)
    assert not missing, missing
