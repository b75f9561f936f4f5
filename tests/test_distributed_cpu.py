"""CPU, world_size 2, gloo: the shard -> per-rank compute -> gather -> reorder -> fold logic of
the one-process-per-GPU path, with the per-rank compute played by the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multiple-sequence-alignment-openmp-openmpi_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nchunks, out_path):
    import importlib

    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dmod = importlib.import_module(PKG + ".distributed")
        workloads = importlib.import_module(PKG + ".workloads")
        pkg = importlib.import_module(PKG)
        from oracle import binding as oracle

        genes = workloads.mixed(9, seed=11, lo=20, hi=140)

        def compute(genes, pxy, pgap, ids):  # stands in for the GPU shard
            pen, dig = [], []
            for t in ids:
                i, j = pkg.task_ij(int(t))
                p, _, _, h = oracle.pair(genes[i], genes[j], pxy, pgap)
                pen.append(p)
                dig.append(np.frombuffer(bytes.fromhex(h), dtype=np.uint8))
            return np.array(pen, dtype=np.int32), np.stack(dig)

        res = dmod.all_pairs_distributed(genes, 3, 2, nchunks=nchunks, compute=compute)
        if rank == 0:
            digest, pen = res
            od, op, _ = oracle.all_pairs(genes, 3, 2)
            ok = digest == od and np.array_equal(pen, op)
            open(out_path, "w").write("ok" if ok else f"mismatch {digest} {od}")
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nchunks", [1, 3])
def test_two_rank_gloo_digest_matches_oracle(tmp_path, nchunks):
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), nchunks, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_shard_plan_covers_all_tasks():
    import importlib

    sys.path.insert(0, ROOT)
    dmod = importlib.import_module(PKG + ".distributed")
    lens = [100, 2000, 350, 900, 1200, 60, 777]
    plan = dmod.ShardPlan(lens, world=3, nchunks=4)
    seen = np.concatenate([plan.rank_ids(r) for r in range(3)])
    assert np.array_equal(np.sort(seen), np.arange(21))
    for c in range(plan.nchunks):
        ids = np.concatenate(plan.ids[c])
        if len(ids):
            assert ids.min() >= plan.bounds[c] and ids.max() < plan.bounds[c + 1]
        assert len(ids) == plan.bounds[c + 1] - plan.bounds[c]
