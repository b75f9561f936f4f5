"""CPU: pins the oracle (oracle/seqalk_oracle.c) against the reference's golden outputs and,
when oracle/_ref exists (build container), against the unmodified reference binaries."""
import hashlib
import os
import random

import numpy as np
import pytest


def test_oracle_sha512_matches_hashlib(oracle):
    rnd = random.Random(5)
    for n in [0, 1, 3, 55, 111, 112, 113, 127, 128, 129, 239, 240, 241, 255, 256, 257, 1000, 100000]:
        data = bytes(rnd.randrange(256) for _ in range(n))
        assert oracle.sha512_hex(data) == hashlib.sha512(data).hexdigest()


@pytest.mark.parametrize("name", ["mseq", "mseq1", "xulin_test"])
def test_oracle_reproduces_golden(oracle, load_fixture, name):
    pxy, pgap, genes, e = load_fixture(name)
    digest, pen, hashes = oracle.all_pairs(genes, pxy, pgap, want_pair_hashes=True)
    assert digest == e["digest"]
    assert list(pen) == e["penalties"]
    assert hashes == e["pair_hashes"]


def test_oracle_big13_last_pairs(oracle, load_fixture):
    """Full big13 takes ~4 min on 8 cores (done once by tests/golden/make_golden.py, where the
    oracle reproduced the reference's digest and all 78 penalties).  Here: the three cheapest
    pairs against the reference's captured penalties and the stored per-pair hashes."""
    pxy, pgap, genes, e = load_fixture("big13")
    for (i, j) in [(12, 11), (12, 10), (11, 10)]:
        tid = i * (i - 1) // 2 + j
        pen, a1, a2, h = oracle.pair(genes[i], genes[j], pxy, pgap)
        assert pen == e["penalties"][tid]
        assert h == e["pair_hashes"][tid]
        assert a1.replace(b"_", b"") == genes[i] and a2.replace(b"_", b"") == genes[j]


def test_oracle_pair_definition(oracle):
    """problemhash = H(H(a1) ++ H(a2)); alignment strings strip back to the inputs."""
    x, y = b"AGGGCT", b"AGGCA"
    pen, a1, a2, h = oracle.pair(x, y, 3, 2)
    assert len(a1) == len(a2)
    assert a1.replace(b"_", b"") == x and a2.replace(b"_", b"") == y
    hh = lambda b: hashlib.sha512(b).hexdigest().encode()
    assert h == hashlib.sha512(hh(a1) + hh(a2)).hexdigest()
    cost = sum(2 if (p == 95 or q == 95) else (0 if p == q else 3) for p, q in zip(a1, a2))
    assert cost == pen


def _random_case(rnd, workloads):
    k = rnd.randint(2, 6)
    alpha = rnd.choice([b"A", b"AC", b"ACGT", b"ACGTacgt", b"xyzXYZ09"])
    genes = [bytes(rnd.choice(alpha) for _ in range(rnd.randint(1, 150))) for _ in range(k)]
    pxy = rnd.choice([0, 1, 2, 3, 5, 9])
    pgap = rnd.choice([0, 1, 2, 3, 7])
    return pxy, pgap, genes


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_seq")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_unmodified_reference_fuzz(oracle, workloads):
    rnd = random.Random(2024)
    for it in range(40):
        pxy, pgap, genes = _random_case(rnd, workloads)
        doc = workloads.format_input(pxy, pgap, genes)
        binary = oracle.REF_SEQ if it % 4 else oracle.REF_PAR
        _, rd, rp = oracle.run_reference(binary, doc, timeout=120)
        od, op, _ = oracle.all_pairs(genes, pxy, pgap)
        assert od == rd, (it, pxy, pgap, genes)
        assert list(op) == rp


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_seq")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_underscore_inputs_match_unmodified_reference(oracle, workloads):
    """'_' inside the inputs: the reference's strip scan (skeleton:135-145) cuts the alignment at the
    LAST column with '_' in both strings, not only at the leading pad."""
    rnd = random.Random(95)
    for it in range(30):
        k = rnd.randint(2, 5)
        alpha = rnd.choice([b"A_", b"AC_", b"ACG_", b"ACGT_", b"__A"])
        genes = [bytes(rnd.choice(alpha) for _ in range(rnd.randint(1, 120))) for _ in range(k)]
        pxy, pgap = rnd.choice([0, 1, 3, 5]), rnd.choice([0, 1, 2, 4])
        doc = workloads.format_input(pxy, pgap, genes)
        _, rd, rp = oracle.run_reference(oracle.REF_SEQ if it % 3 else oracle.REF_PAR, doc, timeout=120)
        od, op, _ = oracle.all_pairs(genes, pxy, pgap)
        assert od == rd, (it, pxy, pgap, genes)
        assert list(op) == rp


def test_oracle_underscore_golden(oracle, underscore_cases):
    for c in underscore_cases:
        od, op, _ = oracle.all_pairs([g.encode("latin-1") for g in c["genes"]], c["pxy"], c["pgap"])
        assert od == c["digest"] and list(op) == c["penalties"]


def test_oracle_empty_sequence_follows_sequential_skeleton(oracle):
    pen, a1, a2, _ = oracle.pair(b"", b"ACGT", 3, 2)
    assert (pen, a1, a2) == (8, b"____", b"ACGT")
    pen, a1, a2, _ = oracle.pair(b"AC", b"", 3, 2)
    assert (pen, a1, a2) == (4, b"AC", b"__")


def test_oracle_negative_penalties_match_reference_semantics(oracle):
    """Negative penalties are legal integers for the reference; the oracle follows the same
    'match takes the diagonal, traceback tests in order' rules (checked against _ref in the
    fuzz above for the non-negative grid; here only self-consistency)."""
    x, y = b"ACGTTGCA", b"ACTTGGCA"
    pen, a1, a2, _ = oracle.pair(x, y, -1, 2)
    assert a1.replace(b"_", b"") == x and a2.replace(b"_", b"") == y
    assert oracle.pair_score(x, y, -1, 2) == pen
