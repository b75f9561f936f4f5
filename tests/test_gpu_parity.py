"""GPU (B200): the CUDA path, called through the C ABI, against the CPU oracle on the same
inputs, against the reference's golden outputs, and — at sizes the oracle cannot reach in
seconds — through size-independent properties.  Everything here is bit-exact (int32 / bytes)."""
import hashlib
import os
import random
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _init(seqalk):
    seqalk.init(1)
    yield


def _check_all_pairs(seqalk, oracle, genes, pxy, pgap):
    digest, pen = seqalk.get_minimum_penalties(genes, pxy, pgap)
    od, op, oh = oracle.all_pairs(genes, pxy, pgap, want_pair_hashes=True)
    if digest != od or not np.array_equal(pen, op):
        # locate the first differing pair for the failure message
        with seqalk.Job(genes, pxy, pgap) as job:
            job.run()
            jp, jd = job.fetch()
        bad = [t for t in range(len(op)) if jp[t] != op[t] or jd[t].tobytes().hex() != oh[t]]
        raise AssertionError(f"pxy={pxy} pgap={pgap} first bad task {bad[:5]} of {len(op)}; "
                             f"lens={[len(g) for g in genes]}")


# ---------------------------------------------------------------- golden vectors

@pytest.mark.parametrize("name", ["mseq", "mseq1", "xulin_test", "xulin"])
def test_golden_small(seqalk, load_fixture, name):
    pxy, pgap, genes, e = load_fixture(name)
    digest, pen = seqalk.get_minimum_penalties(genes, pxy, pgap)
    assert list(pen) == e["penalties"]
    assert digest == e["digest"]


def test_golden_pair_hashes(seqalk, load_fixture):
    pxy, pgap, genes, e = load_fixture("xulin_test")
    with seqalk.Job(genes, pxy, pgap) as job:
        job.run()
        pen, dig = job.fetch()
    assert [d.tobytes().hex() for d in dig] == e["pair_hashes"]
    assert list(pen) == e["penalties"]
    assert seqalk.fold_chain(dig) == e["digest"]


@pytest.mark.parametrize("name", ["big13", "big13_2"])
def test_golden_big13(seqalk, load_fixture, name):
    """BASELINE config 2: 278.5 G cells, bit-exact digest vs the reference's CPU run."""
    pxy, pgap, genes, e = load_fixture(name)
    digest, pen = seqalk.get_minimum_penalties(genes, pxy, pgap)
    assert list(pen) == e["penalties"]
    assert digest == e["digest"]


def test_driver_stdin_stdout_bytes(load_fixture, workloads):
    """seqalkway: same stdout bytes as the reference main, except the time value."""
    exe = os.path.join(ROOT, "multiple-sequence-alignment-openmp-openmpi_b200", "seqalkway")
    for name in ["mseq", "mseq1"]:
        pxy, pgap, genes, e = load_fixture(name)
        raw = open(os.path.join(ROOT, "tests", "golden", e["file"]), "rb").read()
        out = subprocess.run([exe], input=raw, stdout=subprocess.PIPE, check=True,
                             env=dict(os.environ, SEQALK_GPUS="1")).stdout.decode()
        lines = out.split("\n")
        assert lines[0].startswith("Time: ") and lines[0].endswith(" us")
        assert "\n".join(lines[1:]) == e["digest"] + "\n" + "".join(f"{p} " for p in e["penalties"]) + "\n"
    # any whitespace separates tokens (operator>> semantics), no trailing newline needed
    pxy, pgap, genes, e = load_fixture("mseq")
    doc = b"\t3 \r\n 2\n\n3  " + b" \t ".join(genes)
    out = subprocess.run([exe], input=doc, stdout=subprocess.PIPE, check=True,
                         env=dict(os.environ, SEQALK_GPUS="1")).stdout.decode().split("\n")
    assert out[1] == e["digest"] and out[2] == "5 4 9 "


# ---------------------------------------------------------------- oracle parity, seeded

def test_fuzz_small_vs_oracle(seqalk, oracle):
    rnd = random.Random(99)
    for it in range(60):
        k = rnd.randint(2, 7)
        alpha = rnd.choice([b"A", b"AC", b"ACGT", b"ACGTacgt", b"xyzXYZ09"])
        genes = [bytes(rnd.choice(alpha) for _ in range(rnd.randint(1, 300))) for _ in range(k)]
        pxy = rnd.choice([0, 1, 2, 3, 5, 9])
        pgap = rnd.choice([0, 1, 2, 3, 7])
        _check_all_pairs(seqalk, oracle, genes, pxy, pgap)


def test_strip_and_macro_column_boundaries(seqalk, oracle):
    """Lengths around the kernel's tiling: 8 rows per lane, 256 rows per strip, 4-column macro
    columns, 16 B direction words."""
    rnd = random.Random(7)
    lens = [1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 255, 256, 257, 511, 512, 513, 767, 769, 1023, 1025]
    genes = [bytes(rnd.choice(b"ACGT") for _ in range(L)) for L in lens]
    _check_all_pairs(seqalk, oracle, genes, 3, 2)


def test_general_kernel_negative_and_large_penalties(seqalk, oracle):
    rnd = random.Random(3)
    genes = [bytes(rnd.choice(b"ACG") for _ in range(rnd.randint(1, 400))) for _ in range(6)]
    for pxy, pgap in [(-1, 2), (3, -1), (-2, -3), (0, 0), (1000000, 3), (2, 900000)]:
        _check_all_pairs(seqalk, oracle, genes, pxy, pgap)


def test_forced_general_kernel_equals_fast_kernel(seqalk, oracle, monkeypatch):
    rnd = random.Random(4)
    genes = [bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(200, 900))) for _ in range(6)]
    d_fast, p_fast = seqalk.get_minimum_penalties(genes, 3, 2)
    monkeypatch.setenv("SEQALK_FORCE_GENERAL", "1")
    d_gen, p_gen = seqalk.get_minimum_penalties(genes, 3, 2)
    assert d_fast == d_gen and np.array_equal(p_fast, p_gen)
    od, op, _ = oracle.all_pairs(genes, 3, 2)
    assert d_fast == od


def test_special_shapes(seqalk, oracle):
    base = bytes(random.Random(1).choice(b"ACGT") for _ in range(700))
    cases = [
        [base, base],                      # equal strings: pure diagonal
        [base, base[100:400]],             # substring: long pure-gap runs (tested in big13)
        [base[100:400], base],
        [b"A" * 500, b"A" * 137],          # one-letter alphabet
        [b"A" * 300, b"C" * 300],          # no match anywhere
        [b"AcGt" * 50, b"ACGT" * 50],      # case-sensitive comparison
        [b"G", b"G"], [b"G", b"T"], [b"G", base], [base, b"T"],
        [bytes([200, 201, 250, 128] * 40), bytes([201, 200, 128, 250] * 37)],  # high bytes
    ]
    for genes in cases:
        _check_all_pairs(seqalk, oracle, genes, 3, 2)
        _check_all_pairs(seqalk, oracle, genes, 5, 1)


def test_single_pair_arrays(seqalk, oracle):
    """getMinimumPenalty: right-aligned int arrays, '_' pad, strip scan gives the strings."""
    rnd = random.Random(12)
    for _ in range(10):
        x = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, 600)))
        y = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, 600)))
        pen, xans, yans = seqalk.get_minimum_penalty(x, y, 3, 2)
        open_, a1, a2, _ = oracle.pair(x, y, 3, 2)
        assert pen == open_
        l, L = len(x) + len(y), len(a1)
        assert bytes(int(v) for v in xans[l - L + 1:]) == a1
        assert bytes(int(v) for v in yans[l - L + 1:]) == a2
        assert np.all(xans[1:l - L + 1] == 95) and np.all(yans[1:l - L + 1] == 95)
        assert seqalk.alignment_strings(xans, yans) == (a1, a2)
    xa, ya = [0] * 12, [0] * 12
    assert seqalk.getMinimumPenalty("AGGGCT", "AGGCA", 3, 2, xa, ya) == 5
    pens = [0] * 3
    assert seqalk.getMinimumPenalties(["AGGGCT", "AGGCA", "AAAGGGCT"], 3, 3, 2, pens).startswith("602d0f604e8f")
    assert pens == [5, 4, 9]


def test_batching_does_not_change_results(seqalk, oracle, workloads, monkeypatch):
    genes = workloads.mixed(14, seed=5, lo=300, hi=2500)
    d1, p1 = seqalk.get_minimum_penalties(genes, 3, 2)
    with seqalk.Job(genes, 3, 2) as job:
        job.run()
        assert job.stats().batches == 1
    monkeypatch.setenv("SEQALK_DIR_BUDGET_GB", "0.0005")  # ~0.5 MB of checkpoints/directions per batch
    with seqalk.Job(genes, 3, 2) as job:
        job.run()
        pen, dig = job.fetch()
        st = job.stats()
    # one persistent fill launch per batch (+ the sentinel rows of a few-pair job's batches, fill16.cu)
    assert st.batches > 5 and st.batches <= st.fill_launches <= 2 * st.batches
    assert st.cells == workloads.total_cells(genes)
    assert seqalk.fold_chain(dig) == d1 and np.array_equal(pen, p1)
    od, op, _ = oracle.all_pairs(genes, 3, 2)
    assert d1 == od and np.array_equal(p1, op)


def test_streaming_fold_with_blocking_launches(oracle, workloads):
    """The public call folds the chain batch by batch while later batches are still to be
    enqueued.  With synchronous launches (CUDA_LAUNCH_BLOCKING=1, or a profiler) a batch is
    finished by the time the host looks at it, so the fold catches up with the enqueue loop:
    it must never look past the batches enqueued in this call (their events are unrecorded or
    carry the previous call's record).  Fresh process: the variable is read at CUDA init."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import importlib, sys\n"
        "sys.path.insert(0, %r)\n"
        "pkg = importlib.import_module('multiple-sequence-alignment-openmp-openmpi_b200')\n"
        "w = importlib.import_module('multiple-sequence-alignment-openmp-openmpi_b200.workloads')\n"
        "for seed in (5, 6):\n"
        "    genes = w.mixed(14, seed=seed, lo=300, hi=2500)\n"
        "    d, p = pkg.get_minimum_penalties(genes, 3, 2)\n"
        "    print(d, ' '.join(str(int(x)) for x in p))\n" % root)
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1", SEQALK_DIR_BUDGET_GB="0.0005")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 2
    for seed, line in zip((5, 6), lines):
        od, op, _ = oracle.all_pairs(workloads.mixed(14, seed=seed, lo=300, hi=2500), 3, 2)
        digest, pens = line.split(" ", 1)
        assert digest == od and [int(x) for x in pens.split()] == [int(x) for x in op]


def test_job_subset_and_device_pointers(seqalk, oracle, workloads):
    import torch

    genes = workloads.mixed(10, seed=8, lo=100, hi=1200)
    od, op, oh = oracle.all_pairs(genes, 3, 2, want_pair_hashes=True)
    ids = seqalk.partition([len(g) for g in genes], 3, 1)
    pen_t = torch.zeros(len(ids), dtype=torch.int32, device="cuda")
    dig_t = torch.zeros((len(ids), 64), dtype=torch.uint8, device="cuda")
    with seqalk.Job(genes, 3, 2, ids, device=0) as job:
        job.run(pen_t.data_ptr(), dig_t.data_ptr())
    torch.cuda.synchronize()
    assert pen_t.cpu().tolist() == [int(op[t]) for t in ids]
    assert [bytes(r).hex() for r in dig_t.cpu().numpy()] == [oh[t] for t in ids]


def test_distributed_single_rank_gpu(seqalk, oracle, workloads):
    import importlib

    dmod = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200.distributed")
    genes = workloads.mixed(12, seed=21, lo=100, hi=900)
    digest, pen = dmod.all_pairs_distributed(genes, 3, 2, nchunks=3, device=0)
    od, op, _ = oracle.all_pairs(genes, 3, 2)
    assert digest == od and np.array_equal(pen, op)


# ---------------------------------------------------------------- BASELINE shapes

def test_cfg3_sample_vs_oracle(seqalk, oracle, workloads):
    """Config 3 shape (10k x 10k, uniform ACGT): a 10-sequence prefix of the seeded input."""
    genes = workloads.uniform(200, 10000, 200)[:10]
    _check_all_pairs(seqalk, oracle, genes, 3, 2)


def test_cfg4_sample_vs_oracle(seqalk, oracle, workloads):
    """Config 4 shape (1k-20k mixed): seeded sample incl. the extremes of length ratio."""
    genes = workloads.mixed(1000, 1000)
    lens = [len(g) for g in genes]
    order = np.argsort(lens)
    pick = sorted(set(int(v) for v in list(order[:3]) + list(order[-3:]) + [0, 1, 2, 500]))
    _check_all_pairs(seqalk, oracle, [genes[i] for i in pick], 3, 2)


def _alignment_properties(x, y, pen, xans, yans, pxy, pgap, seqalk):
    a1, a2 = seqalk.alignment_strings(xans, yans)
    A1, A2 = np.frombuffer(a1, np.uint8), np.frombuffer(a2, np.uint8)
    assert len(A1) == len(A2)
    assert A1[A1 != 95].tobytes() == x and A2[A2 != 95].tobytes() == y
    assert not np.any((A1 == 95) & (A2 == 95))
    gaps = int(np.sum((A1 == 95) | (A2 == 95)))
    mism = int(np.sum((A1 != 95) & (A2 != 95) & (A1 != A2)))
    assert gaps * pgap + mism * pxy == pen           # the path costs exactly dp[m][n]
    return hashlib.sha512((hashlib.sha512(a1).hexdigest() + hashlib.sha512(a2).hexdigest()).encode()).hexdigest()


def test_cfg5_full_size_properties(seqalk, oracle, workloads):
    """Config 5: one 200 000 x 200 000 pair (4e10 cells, 10 GB of packed directions).  The
    oracle needs minutes for this, so: alignment strips back to x / y, its cost equals the
    reported penalty, the device digest equals hashlib over the returned strings, the score is
    symmetric, and a 20 000-column prefix problem matches the oracle's linear-memory score."""
    x, y = workloads.uniform(2, 200000, 2)[::-1]  # task (1,0): x = genes[1], y = genes[0]
    # the 10 GB packed-direction path (dirs mode) and the default tile mode must agree
    os.environ["SEQALK_TRACEBACK"] = "dirs"
    try:
        pen_d, xans_d, yans_d = seqalk.get_minimum_penalty(x, y, 3, 2)
    finally:
        os.environ.pop("SEQALK_TRACEBACK", None)
    pen, xans, yans = seqalk.get_minimum_penalty(x, y, 3, 2)
    assert pen_d == pen and np.array_equal(xans_d, xans) and np.array_equal(yans_d, yans)
    ph = _alignment_properties(x, y, pen, xans, yans, 3, 2, seqalk)
    digest, pens = seqalk.get_minimum_penalties([y, x], 3, 2)
    assert pens[0] == pen
    assert digest == hashlib.sha512(ph.encode()).hexdigest()
    pen_t, _, _ = seqalk.get_minimum_penalty(y, x, 3, 2)
    assert pen_t == pen
    p2, _, _ = seqalk.get_minimum_penalty(x[:20000], y[:20000], 3, 2)
    assert p2 == oracle.pair_score(x[:20000], y[:20000], 3, 2)


@pytest.mark.parametrize("env", [
    {"SEQALK_FILL_R": "8", "SEQALK_FILL_TAB": "1", "SEQALK_FILL_ADD": "0"},
    {"SEQALK_FILL_R": "8", "SEQALK_FILL_TAB": "0", "SEQALK_FILL_ADD": "0"},
    {"SEQALK_FILL_R": "8", "SEQALK_FILL_TAB": "1", "SEQALK_FILL_ADD": "1"},
    {"SEQALK_FILL_R": "8", "SEQALK_FILL_TAB": "0", "SEQALK_FILL_ADD": "1"},
    {"SEQALK_FILL_R": "16", "SEQALK_FILL_TAB": "1", "SEQALK_FILL_ADD": "0"},
    {"SEQALK_FILL_R": "16", "SEQALK_FILL_TAB": "0", "SEQALK_FILL_ADD": "1"},
    {"SEQALK_FILL_R": "16", "SEQALK_FORCE_GENERAL": "1"},
    {"SEQALK_FILL_R": "8", "SEQALK_SLOTS": "1"},
    {"SEQALK_TRACEBACK": "dirs"},
    {"SEQALK_TRACEBACK": "dirs", "SEQALK_FILL_ADD": "2"},
    {"SEQALK_TRACEBACK": "tiles", "SEQALK_FILL_ADD": "2"},
    {"SEQALK_TRACEBACK": "tiles", "SEQALK_FILL_R": "16"},
    {"SEQALK_TRACEBACK": "tiles", "SEQALK_FILL_R": "16", "SEQALK_FILL_TAB": "0"},
    {"SEQALK_TRACEBACK": "tiles", "SEQALK_FILL_R": "16", "SEQALK_FORCE_GENERAL": "1"},
    {"SEQALK_TRACEBACK": "dirs", "SEQALK_FILL_R": "8", "SEQALK_FILL_TAB": "0"},
    {"SEQALK_TRACEBACK": "tiles"},
    {"SEQALK_TRACEBACK": "tiles", "SEQALK_FILL_TAB": "0"},
    {"SEQALK_TRACEBACK": "tiles", "SEQALK_FILL_TAB": "0", "SEQALK_FILL_ADD": "1"},
    {"SEQALK_TRACEBACK": "tiles", "SEQALK_FORCE_GENERAL": "1"},
    {"SEQALK_TRACEBACK": "tiles", "SEQALK_DIR_BUDGET_GB": "0.001"},
    # the many-pair production path on this small job: packed fill (one persistent launch over all
    # strip pairs) + half-warp tile traceback, and the alternatives kept behind knobs
    {"SEQALK_FILL_R": "32", "SEQALK_TB_LONG": "0"},
    {"SEQALK_FILL_R": "32", "SEQALK_TB_LONG": "0", "SEQALK_TB_HALF": "0"},
    {"SEQALK_FILL_R": "32", "SEQALK_TB_LONG": "0", "SEQALK_FILL16_SPLIT": "1", "SEQALK_QUEUE_REMAINING": "0"},
    {"SEQALK_FILL_R": "32", "SEQALK_TB_LONG": "0", "SEQALK_FILL16_RP": "16"},
    {"SEQALK_FILL_R": "32", "SEQALK_TB_LONG": "0", "SEQALK_DIR_BUDGET_GB": "0.001"},
    {"SEQALK_FILL_R": "32", "SEQALK_TB_LONG": "0", "SEQALK_DIR_BUDGET_GB": "0.001", "SEQALK_SLOTS": "2",
     "SEQALK_POST_PRIORITY": "0"},
    # 32-bit fill (int32 column checkpoints every 32 macro columns) under the many-pair tracebacks
    {"SEQALK_FILL16": "0", "SEQALK_TB_LONG": "0"},
    {"SEQALK_FILL16": "0", "SEQALK_TB_LONG": "0", "SEQALK_TB_HALF": "0"},
    {"SEQALK_FILL16": "0", "SEQALK_TB_LONG": "0", "SEQALK_FORCE_GENERAL": "1"},
    # banded traceback in chapters (the path of a few LONG pairs), forced onto this small job: default
    # chapter length, chapters of 4 levels with a band of 3 tiles, chapters of ONE level without
    # neighbours (most tiles are recomputed by the walking warp), every fill under it, several batches;
    # per-pair hashes computed by the host
    {"SEQALK_TB_BAND": "1"},
    {"SEQALK_TB_BAND": "1", "SEQALK_TB_BAND_LEVELS": "3", "SEQALK_TB_BAND_HALFWIDTH": "1"},
    {"SEQALK_TB_BAND": "1", "SEQALK_TB_BAND_LEVELS": "0", "SEQALK_TB_BAND_HALFWIDTH": "0"},
    {"SEQALK_TB_BAND": "1", "SEQALK_TB_BAND_LEVELS": "5", "SEQALK_FILL16": "0"},
    {"SEQALK_TB_BAND": "1", "SEQALK_TB_BAND_LEVELS": "2", "SEQALK_FORCE_GENERAL": "1"},
    {"SEQALK_TB_BAND": "1", "SEQALK_TB_BAND_LEVELS": "7", "SEQALK_FILL_TAB": "0", "SEQALK_DIR_BUDGET_GB": "0.001"},
    {"SEQALK_HOST_HASH": "1"},
    # thin strip pairs of the packed fill (512 / 256 rows: the few-long-pairs geometry) under every traceback
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "8"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "4"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "8", "SEQALK_TB_LONG": "0"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "4", "SEQALK_TB_LONG": "0", "SEQALK_DIR_BUDGET_GB": "0.001"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "4", "SEQALK_TB_BAND": "1", "SEQALK_TB_BAND_LEVELS": "3"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "8", "SEQALK_TB_BAND": "1", "SEQALK_TB_BAND_CHAPTERS": "2"},
    # packed fill builds: column checkpoints by the one-lane branch instead of predicated stores through a
    # running pointer (the default), register budgets, pipelined lane hand-off, resident CTAs per SM
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_PREDCK": "0", "SEQALK_TB_LONG": "0"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_PREDCK": "0", "SEQALK_FILL_MINB": "3", "SEQALK_TB_LONG": "0"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL_MINB": "4", "SEQALK_TB_LONG": "0", "SEQALK_DIR_BUDGET_GB": "0.001"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "16", "SEQALK_FILL_MINB": "6"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "16", "SEQALK_FILL16_PREDCK": "0", "SEQALK_FILL_CTAS_PER_SM": "1"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "8", "SEQALK_FILL16_PIPE": "0"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_SENT": "0"},  # progress counters instead of sentinel hand-off rows
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_SENT": "1", "SEQALK_DIR_BUDGET_GB": "0.001", "SEQALK_SLOTS": "2"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "8", "SEQALK_FILL16_PIPE": "1", "SEQALK_FILL_MINB": "8"},
    {"SEQALK_FILL_R": "32", "SEQALK_FILL16_RP": "8", "SEQALK_FILL16_PREDCK": "0", "SEQALK_FILL16_PIPE": "0",
     "SEQALK_DIR_BUDGET_GB": "0.001"},
], ids=lambda e: ",".join(f"{k[7:]}={v}" for k, v in e.items()))
def test_every_kernel_variant_is_bit_exact(seqalk, oracle, workloads, monkeypatch, env):
    """Rows per lane (8/16), cost by compare or by 4-letter table, adds on ALU or FMA pipe,
    general (any-integer) kernel: all must reproduce the oracle bit for bit."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rnd = random.Random(17)
    lens = [1, 5, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1023, 1025, 1030, 1536, 1537, 2049, 2560]
    genes = [bytes(rnd.choice(b"ACGT") for _ in range(L)) for L in lens]
    with seqalk.Job(genes, 3, 2) as job:  # a fresh worker picks the env up
        job.run()
        pen, dig = job.fetch()
        if env.get("SEQALK_FILL_R") == "32":  # the packed fill really ran
            assert job.stats().fill_packed16 == 1
    od, op, oh = oracle.all_pairs(genes, 3, 2, want_pair_hashes=True)
    assert np.array_equal(pen, op)
    assert [d.tobytes().hex() for d in dig] == oh
    genes3 = [bytes(rnd.choice(b"ACG") for _ in range(rnd.randint(1, 700))) for _ in range(5)]
    with seqalk.Job(genes3, 5, 1) as job:
        job.run()
        pen, dig = job.fetch()
    od, op, oh = oracle.all_pairs(genes3, 5, 1, want_pair_hashes=True)
    assert np.array_equal(pen, op) and [d.tobytes().hex() for d in dig] == oh
    # a 20-letter alphabet: the compare kernels (no cost table), whatever traceback the env selects
    genes20 = [bytes(rnd.choice(b"ACDEFGHIKLMNPQRSTVWY") for _ in range(rnd.randint(100, 900))) for _ in range(7)]
    with seqalk.Job(genes20, 4, 3) as job:
        job.run()
        pen, dig = job.fetch()
    od, op, oh = oracle.all_pairs(genes20, 4, 3, want_pair_hashes=True)
    assert np.array_equal(pen, op) and [d.tobytes().hex() for d in dig] == oh


def _gpu_count():
    import torch

    return torch.cuda.device_count()


def test_two_gpus_in_process_match_one_gpu(seqalk, oracle, workloads):
    """seqalk_init(2): pairs sharded by cells over two GPUs from one process, results stored
    by the digest kernel straight into GPU 0's task-ordered arrays (peer stores over NVLink)."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    genes = workloads.mixed(16, seed=31, lo=200, hi=3000)
    od, op, _ = oracle.all_pairs(genes, 3, 2)
    try:
        seqalk.init(2)
        assert seqalk.device_count() == 2
        d2, p2 = seqalk.get_minimum_penalties(genes, 3, 2)
        os.environ["SEQALK_P2P"] = "0"   # unfused comparison: per-GPU D2H + host scatter
        seqalk.init(2)
        d2b, p2b = seqalk.get_minimum_penalties(genes, 3, 2)
    finally:
        os.environ.pop("SEQALK_P2P", None)
        seqalk.init(1)
    assert d2 == od and np.array_equal(p2, op)
    assert d2b == od and np.array_equal(p2b, op)


def test_two_gpus_in_process_chunked_fold(seqalk, oracle):
    """79 800 pairs on two GPUs of one process: the job is cut into two chunks of task ids,
    each balanced over the GPUs; the host folds chunk 0 while the GPUs run chunk 1."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    rnd = random.Random(8)
    genes = [bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, 40))) for _ in range(400)]
    od, op, _ = oracle.all_pairs(genes, 3, 2)
    try:
        seqalk.init(2)
        d2, p2 = seqalk.get_minimum_penalties(genes, 3, 2)
    finally:
        seqalk.init(1)
    assert d2 == od and np.array_equal(p2, op)


@pytest.mark.parametrize("mode", ["tiles", "dirs"])
def test_traceback_modes_on_long_and_skewed_pairs(seqalk, oracle, workloads, monkeypatch, mode):
    """Both traceback modes (directions in HBM / per-tile recompute from checkpoints) on pairs
    that cross many 256-row and 256-column tile edges, incl. very unequal lengths and a
    substring pair (long pure-gap runs along a tile edge)."""
    monkeypatch.setenv("SEQALK_TRACEBACK", mode)
    rnd = random.Random(23)
    base = bytes(rnd.choice(b"ACGT") for _ in range(6000))
    genes = [base, base[700:5200], bytes(rnd.choice(b"ACGT") for _ in range(5000)),
             bytes(rnd.choice(b"ACGT") for _ in range(300)), bytes(rnd.choice(b"ACGT") for _ in range(2049)),
             bytes(rnd.choice(b"ACGT") for _ in range(257)), base[::-1][:3000]]
    with seqalk.Job(genes, 3, 2) as job:
        job.run()
        pen, dig = job.fetch()
        assert job.stats().tile_mode == (1 if mode == "tiles" else 0)
    od, op, oh = oracle.all_pairs(genes, 3, 2, want_pair_hashes=True)
    assert np.array_equal(pen, op)
    assert [d.tobytes().hex() for d in dig] == oh
    # one long pair through getMinimumPenalty (arrays) in this mode
    x, y = genes[0], genes[2]
    pen1, xans, yans = seqalk.get_minimum_penalty(x, y, 3, 2)
    open_, a1, a2, _ = oracle.pair(x, y, 3, 2)
    assert pen1 == open_ and seqalk.alignment_strings(xans, yans) == (a1, a2)


@pytest.mark.parametrize("levels,halfwidth", [(63, 4), (8, 1), (1, 0), (200, 2)])
def test_banded_traceback_on_long_and_skewed_pairs(seqalk, oracle, monkeypatch, levels, halfwidth):
    """The banded traceback (produce the tiles near the straight line to the origin, then walk) on pairs
    whose path does NOT follow that line — very unequal lengths, a substring pair (the path runs along
    the diagonal and ends in one long gap) — so that part of the tiles come from the band and part are
    recomputed by the walking warp; chapter lengths from 2 levels to the whole pair."""
    monkeypatch.setenv("SEQALK_TB_BAND", "1")
    monkeypatch.setenv("SEQALK_TB_BAND_LEVELS", str(levels))
    monkeypatch.setenv("SEQALK_TB_BAND_HALFWIDTH", str(halfwidth))
    rnd = random.Random(29)
    base = bytes(rnd.choice(b"ACGT") for _ in range(9000))
    genes = [base, base[1200:7700], bytes(rnd.choice(b"ACGT") for _ in range(7000)),
             bytes(rnd.choice(b"ACGT") for _ in range(300)), bytes(rnd.choice(b"ACGT") for _ in range(2049)),
             base[::-1][:3000], b"A" * 1500]
    with seqalk.Job(genes, 3, 2) as job:
        job.run()
        pen, dig = job.fetch()
        assert job.stats().tile_mode == 1
    od, op, oh = oracle.all_pairs(genes, 3, 2, want_pair_hashes=True)
    assert np.array_equal(pen, op)
    assert [d.tobytes().hex() for d in dig] == oh
    x, y = genes[0], genes[2]
    pen1, xans, yans = seqalk.get_minimum_penalty(x, y, 3, 2)
    open_, a1, a2, _ = oracle.pair(x, y, 3, 2)
    assert pen1 == open_ and seqalk.alignment_strings(xans, yans) == (a1, a2)


def test_cfg4_random_task_sample_vs_oracle(seqalk, oracle, workloads):
    """Config 4 at full shape: 48 seeded task ids out of the 499 500 (plus the largest and the
    most skewed pair), run as a job on the full k=1000 input, each checked against the oracle's
    penalty and problemhash.  (The full digest is checked for equality across traceback modes
    and GPU counts by bench.py / profiles; the CPU cannot do 5.5e13 cells.)"""
    from concurrent.futures import ThreadPoolExecutor

    genes = workloads.mixed(1000, 1000)
    lens = np.array([len(g) for g in genes])
    rnd = random.Random(4)
    T = seqalk.num_tasks(1000)
    ids = sorted(set(rnd.randrange(T) for _ in range(48)))
    big = np.argsort(lens)[-2:]
    i, j = int(max(big)), int(min(big))
    ids.append(i * (i - 1) // 2 + j)                      # largest pair
    lo, hi = int(np.argmin(lens)), int(np.argmax(lens))
    i, j = max(lo, hi), min(lo, hi)
    ids.append(i * (i - 1) // 2 + j)                      # most skewed pair
    ids = np.array(sorted(set(ids)), dtype=np.int64)
    with seqalk.Job(genes, 3, 2, ids) as job:
        job.run()
        pen, dig = job.fetch()

    def check(k):
        i, j = seqalk.task_ij(int(ids[k]))
        p, _, _, h = oracle.pair(genes[i], genes[j], 3, 2)
        return p == pen[k] and h == dig[k].tobytes().hex()

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        ok = list(ex.map(check, range(len(ids))))
    assert all(ok), [int(ids[k]) for k, v in enumerate(ok) if not v]


def test_job_chunks_and_async_hand_off(seqalk, oracle, workloads):
    """seqalk_job_set_chunks / run_chunk / start_async + wait_chunk: same results as one run."""
    import torch

    genes = workloads.mixed(14, seed=3, lo=150, hi=2000)
    od, op, oh = oracle.all_pairs(genes, 3, 2, want_pair_hashes=True)
    T = len(op)
    ends = [10, 10, 37, T]  # includes an empty chunk
    with seqalk.Job(genes, 3, 2) as job:
        job.set_chunks(ends)
        for c in range(len(ends)):
            job.run_chunk(c)
        pen, dig = job.fetch()
        assert np.array_equal(pen, op) and [d.tobytes().hex() for d in dig] == oh
        pen_t = torch.zeros(T, dtype=torch.int32, device="cuda")
        dig_t = torch.zeros((T, 64), dtype=torch.uint8, device="cuda")
        job.start_async(pen_t.data_ptr(), dig_t.data_ptr())
        lo = 0
        for c, hi in enumerate(ends):
            job.wait_chunk(c, torch.cuda.current_stream().cuda_stream)
            part = pen_t[lo:hi].cpu().numpy()  # ordered after the chunk by the stream wait
            assert np.array_equal(part, op[lo:hi])
            lo = hi
        job.finish_async()
        assert [bytes(r).hex() for r in dig_t.cpu().numpy()] == oh


def test_many_tiny_sequences(seqalk, oracle):
    """k = 400 sequences of 1..40 letters: 79 800 pairs, most of them far smaller than one strip."""
    rnd = random.Random(8)
    genes = [bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, 40))) for _ in range(400)]
    digest, pen = seqalk.get_minimum_penalties(genes, 3, 2)
    od, op, _ = oracle.all_pairs(genes, 3, 2)
    assert digest == od and np.array_equal(pen, op)
