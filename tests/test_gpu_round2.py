"""GPU (B200), round 2: full-size oracle goldens of BASELINE configs 3-5, the score-only entry
points and the on-device fill/traceback cross-check, inputs containing '_', empty sequences, the
error paths of the C ABI, and the literal drop-in (reference main + seqalk.hpp)."""
import ctypes
import hashlib
import os
import random
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")


@pytest.fixture(scope="module", autouse=True)
def _init(seqalk):
    seqalk.init(1)
    yield


# ---------------------------------------------------------------- full-size goldens (oracle, CPU container)

def test_cfg5_full_size_golden(seqalk, workloads, fullsize):
    """Config 5: the single 200 000 x 200 000 pair against the oracle's penalty, problemhash and
    alignment hashes (tests/golden/make_golden_fullsize.py; seqalign-mpi-skeleton.cpp:117-175)."""
    g = fullsize["cfg5"]
    pxy, pgap, genes = workloads.make("cfg5")
    digest, pen = seqalk.get_minimum_penalties(genes, pxy, pgap)
    assert int(pen[0]) == g["penalty"]
    assert digest == g["digest"]
    with seqalk.Job(genes, pxy, pgap) as job:
        job.run()
        p, d = job.fetch()
        assert job.stats().self_checked == 1
    assert d[0].tobytes().hex() == g["problemhash"] and int(p[0]) == g["penalty"]
    pen1, xans, yans = seqalk.get_minimum_penalty(genes[1], genes[0], pxy, pgap)
    a1, a2 = seqalk.alignment_strings(xans, yans)
    assert len(a1) == g["alignment_length"]
    assert hashlib.sha512(a1).hexdigest() == g["align1_sha512"]
    assert hashlib.sha512(a2).hexdigest() == g["align2_sha512"]
    assert seqalk.pair_score(genes[1], genes[0], pxy, pgap) == g["penalty"]


def test_cfg4_stratified_sample_golden(seqalk, workloads, fullsize):
    """Config 4: 2 027 of the 499 500 tasks (cell-count deciles x 190, the 64 most skewed, the 32
    largest and 32 smallest), run as one job on the full k = 1000 input: penalty and problemhash of
    every task against the oracle."""
    g = fullsize["cfg4"]
    pxy, pgap, genes = workloads.make("cfg4")
    assert workloads.total_cells(genes) == g["cells"]
    ids = np.asarray(g["task_ids"], dtype=np.int64)
    assert len(ids) >= 2000
    with seqalk.Job(genes, pxy, pgap, ids) as job:
        job.run()
        pen, dig = job.fetch()
        assert job.stats().self_checked == 1
        job.run_scores()
        scores = job.fetch_penalties()
    bad = [int(ids[t]) for t in range(len(ids))
           if int(pen[t]) != g["penalties"][t] or dig[t].tobytes().hex() != g["problemhashes"][t]]
    assert not bad, bad[:10]
    assert np.array_equal(scores, np.asarray(g["penalties"], dtype=np.int32))


def test_cfg3_full_size_golden(seqalk, workloads, fullsize):
    """Config 3: the whole k = 200 x 10 000 job (1.99e12 cells): digest, all 19 900 penalties and
    every problemhash (through their SHA-512) against the oracle."""
    if "cfg3" not in fullsize:
        pytest.skip("cfg3 golden not generated yet")
    g = fullsize["cfg3"]
    pxy, pgap, genes = workloads.make("cfg3")
    digest, pen = seqalk.get_minimum_penalties(genes, pxy, pgap)
    assert digest == g["digest"]
    assert [int(v) for v in pen] == g["penalties"]
    with seqalk.Job(genes, pxy, pgap) as job:
        job.run()
        _, dig = job.fetch()
    hashes = [d.tobytes().hex() for d in dig]
    assert hashes[::50] == g["problemhash_every_50"]
    assert hashlib.sha512("".join(hashes).encode()).hexdigest() == g["problemhashes_sha512"]
    assert np.array_equal(seqalk.all_pairs_scores(genes, pxy, pgap), np.asarray(g["penalties"], dtype=np.int32))


# ---------------------------------------------------------------- score-only entry + self-check

def test_score_only_matches_oracle(seqalk, oracle):
    rnd = random.Random(41)
    for it in range(25):
        k = rnd.randint(2, 7)
        alpha = rnd.choice([b"A", b"AC", b"ACGT", b"ACGTacgt", b"xyzXYZ09"])  # table and compare kernels
        genes = [bytes(rnd.choice(alpha) for _ in range(rnd.randint(1, 1500))) for _ in range(k)]
        pxy, pgap = rnd.choice([0, 1, 2, 3, 5, 9]), rnd.choice([0, 1, 2, 3, 7])
        _, op, _ = oracle.all_pairs(genes, pxy, pgap)
        assert np.array_equal(seqalk.all_pairs_scores(genes, pxy, pgap), op), (it, pxy, pgap)
        assert seqalk.pair_score(genes[1], genes[0], pxy, pgap) == int(op[0])
    # negative penalties: the general kernel (path cost, no corner)
    genes = [bytes(rnd.choice(b"ACG") for _ in range(rnd.randint(1, 400))) for _ in range(5)]
    for pxy, pgap in [(-1, 2), (3, -1), (1000000, 3)]:
        _, op, _ = oracle.all_pairs(genes, pxy, pgap)
        assert np.array_equal(seqalk.all_pairs_scores(genes, pxy, pgap), op)


@pytest.mark.parametrize("mode", ["tiles", "dirs"])
def test_self_check_runs_in_table_mode(seqalk, oracle, workloads, monkeypatch, mode):
    """Every penalty of a DNA job is compared on the device with the fill's own dp[m][n]
    (bottom row of the last strip, padded rows copy row m); lengths around the strip / tile edges."""
    monkeypatch.setenv("SEQALK_TRACEBACK", mode)
    rnd = random.Random(5)
    lens = [1, 3, 4, 5, 127, 128, 129, 511, 512, 513, 1023, 1024, 1025, 1537, 2048, 2049, 3000]
    genes = [bytes(rnd.choice(b"ACGT") for _ in range(L)) for L in lens]
    with seqalk.Job(genes, 3, 2) as job:
        job.run()
        pen, _ = job.fetch()
        assert job.stats().self_checked == 1
        job.run_scores()
        assert np.array_equal(job.fetch_penalties(), pen)
    _, op, _ = oracle.all_pairs(genes, 3, 2)
    assert np.array_equal(pen, op)
    with seqalk.Job([b"xyzXYZ09" * 20, b"zyx90" * 30], 3, 2) as job:  # > 4 letters: no corner, no check
        job.run()
        assert job.stats().self_checked == 0


# ---------------------------------------------------------------- '_' inside the inputs, empty sequences

@pytest.mark.parametrize("mode", ["tiles", "dirs"])
def test_underscore_inputs_follow_the_reference_strip_scan(seqalk, underscore_cases, monkeypatch, mode):
    """Outputs of the UNMODIFIED reference (oracle/_ref/ref_seq, stored in tests/golden/underscore.json)
    for inputs containing '_': the alignment is cut after the last '_'/'_' column (:191-202)."""
    monkeypatch.setenv("SEQALK_TRACEBACK", mode)
    for c in underscore_cases:
        genes = [g.encode("latin-1") for g in c["genes"]]
        with seqalk.Job(genes, c["pxy"], c["pgap"]) as job:  # a fresh worker picks the env up
            job.run()
            pen, dig = job.fetch()
        assert [int(v) for v in pen] == c["penalties"]
        assert seqalk.fold_chain(dig) == c["digest"]


def test_empty_sequences(seqalk, oracle):
    """m = 0 or n = 0 (sequential twin getMinimumPenalty :570-663: penalty (m+n)*pgap, pure tails)."""
    assert seqalk.get_minimum_penalty(b"", b"ACGT", 3, 2)[0] == 8
    pen, xans, yans = seqalk.get_minimum_penalty(b"ACG", b"", 3, 2)
    assert pen == 6 and seqalk.alignment_strings(xans, yans) == (b"ACG", b"___")
    assert seqalk.pair_score(b"", b"", 3, 2) == 0
    genes = [b"ACGT", b"", b"GGA", b"", b"T" * 700]
    digest, pen = seqalk.get_minimum_penalties(genes, 3, 2)
    od, op, _ = oracle.all_pairs(genes, 3, 2)
    assert digest == od and np.array_equal(pen, op)
    assert np.array_equal(seqalk.all_pairs_scores(genes, 3, 2), op)


# ---------------------------------------------------------------- error paths of the C ABI

def test_c_abi_error_paths(seqalk):
    lib = seqalk.lib()
    err = lambda: lib.seqalk_last_error().decode()
    dig = ctypes.create_string_buffer(b"x" * 128, 129)
    pen = (ctypes.c_int32 * 4)()
    a = seqalk._SeqArgs([b"ACGT", b"AGT"])
    # k < 2: no pairs — empty chain state, no penalties, success (reference :237, :282-287)
    assert lib.seqalk_all_pairs(a.ptrs, a.lens_p, 1, 3, 2, pen, dig) == 0 and dig.value == b""
    assert lib.seqalk_all_pairs(None, None, 0, 3, 2, None, dig) == 0 and dig.value == b""
    d1, p1 = seqalk.get_minimum_penalties([b"ACGT"], 3, 2)
    assert d1 == "" and len(p1) == 0
    assert lib.seqalk_all_pairs(a.ptrs, a.lens_p, -1, 3, 2, pen, dig) != 0 and "negative" in err()
    # NULL arguments
    assert lib.seqalk_all_pairs(None, a.lens_p, 2, 3, 2, pen, dig) != 0 and "null" in err()
    assert lib.seqalk_all_pairs(a.ptrs, a.lens_p, 2, 3, 2, None, dig) != 0 and "null" in err()
    assert lib.seqalk_all_pairs(a.ptrs, a.lens_p, 2, 3, 2, pen, None) != 0 and "null" in err()
    assert lib.seqalk_pair(None, 4, a.ptrs[1], 3, 3, 2, None, None, pen) != 0 and "null" in err()
    assert lib.seqalk_pair_score(a.ptrs[0], 4, a.ptrs[1], 3, 3, 2, None) != 0 and "null" in err()
    # negative length; input beyond the 4 GiB the 32-bit offsets address (checked before any byte is read)
    bad = (ctypes.c_int32 * 2)(4, -3)
    assert lib.seqalk_all_pairs(a.ptrs, bad, 2, 3, 2, pen, dig) != 0 and "negative" in err()
    b3 = seqalk._SeqArgs([b"A", b"C", b"G"])
    huge = (ctypes.c_int32 * 3)(2**31 - 1, 2**31 - 1, 2**31 - 1)
    assert lib.seqalk_all_pairs(b3.ptrs, huge, 3, 3, 2, pen, dig) != 0 and "4 GiB" in err()
    # a job whose results went to caller memory cannot be fetched; bad ids / device are rejected
    import torch

    genes = [b"ACGTACGT", b"ACGGT", b"TTGA"]
    pen_t = torch.zeros(3, dtype=torch.int32, device="cuda")
    dig_t = torch.zeros((3, 64), dtype=torch.uint8, device="cuda")
    with seqalk.Job(genes, 3, 2) as job:
        with pytest.raises(seqalk.SeqalkError, match="did not run|caller"):
            job.fetch()
        job.run(pen_t.data_ptr(), dig_t.data_ptr())
        with pytest.raises(seqalk.SeqalkError, match="caller-supplied"):
            job.fetch()
        job.run()
        job.fetch()
        job.run_scores()
        with pytest.raises(seqalk.SeqalkError, match="produced none"):
            job.fetch()
    with pytest.raises(seqalk.SeqalkError, match="task id out of range"):
        seqalk.Job(genes, 3, 2, np.array([0, 3], dtype=np.int64))
    with pytest.raises(seqalk.SeqalkError, match="device index"):
        seqalk.Job(genes, 3, 2, device=99)
    # the library still works after the failures
    assert seqalk.get_minimum_penalties([b"AGGGCT", b"AGGCA", b"AAAGGGCT"], 3, 2)[0].startswith("602d0f604e8f")


# ---------------------------------------------------------------- in-process multi-GPU (fused peer-store gather)

def test_two_gpus_in_process_many_chunks_and_empty_shares(seqalk, oracle, workloads, monkeypatch):
    """seqalk_init(2) + seqalk_all_pairs with the job cut into several chunks of task ids
    (SEQALK_PEER_CHUNKS): start_async / wait_chunk per GPU, peer stores into GPU 0's task-ordered
    arrays, fold of chunk c under chunk c+1 — including chunks of ONE task, where one GPU's share
    of the chunk is empty."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    genes = workloads.mixed(9, seed=12, lo=200, hi=2500)  # 36 tasks
    od, op, _ = oracle.all_pairs(genes, 3, 2)
    try:
        seqalk.init(2)
        for nch in ("5", "36"):
            monkeypatch.setenv("SEQALK_PEER_CHUNKS", nch)
            d2, p2 = seqalk.get_minimum_penalties(genes, 3, 2)
            assert d2 == od and np.array_equal(p2, op), nch
        assert np.array_equal(seqalk.all_pairs_scores(genes, 3, 2), op)
    finally:
        seqalk.init(1)


# ---------------------------------------------------------------- the literal drop-in

@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/dropin not built (needs /root/reference at build time)")
def test_literal_dropin_reference_main_plus_seqalk_hpp(load_fixture):
    """oracle/_ref/dropin = lines 1-82 of the reference's submit/xuliny-seqalkway.cpp (its frozen
    main, :35-77) + `#include "seqalk.hpp"` + the single-rank mpi.h shim, linked against
    libseqalk.so (oracle/Makefile dropin): stdout after the `Time:` line must equal the reference's."""
    for name in ["mseq", "mseq1", "xulin_test", "xulin"]:
        pxy, pgap, genes, e = load_fixture(name)
        raw = open(os.path.join(ROOT, "tests", "golden", e["file"]), "rb").read()
        out = subprocess.run([DROPIN], input=raw, stdout=subprocess.PIPE, check=True, timeout=300,
                             env=dict(os.environ, SEQALK_GPUS="1")).stdout.decode()
        lines = out.split("\n")
        assert lines[0].startswith("Time: ") and lines[0].endswith(" us")
        assert "\n".join(lines[1:]) == e["digest"] + "\n" + "".join(f"{p} " for p in e["penalties"]) + "\n"
    out = subprocess.run([DROPIN], input=b"3 2 1 ACGT", stdout=subprocess.PIPE, check=True, timeout=300).stdout.decode()
    assert out.split("\n")[1:] == ["", "", ""]  # k = 1: empty digest line, empty penalty line
