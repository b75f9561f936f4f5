"""GPU, 2 ranks (one process per GPU; records gathered through pinned host memory + gloo, and
through NCCL for comparison): digest identical to the oracle and to 1 GPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
PKG = "multiple-sequence-alignment-openmp-openmpi_b200"
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dmod = importlib.import_module(PKG + ".distributed")
workloads = importlib.import_module(PKG + ".workloads")
genes = workloads.mixed(20, seed=77, lo=300, hi=4000)
res = dmod.all_pairs_distributed(genes, 3, 2, nchunks=3, device=local)            # host gather (gloo)
res2 = dmod.all_pairs_distributed(genes, 3, 2, nchunks=2, device=local, gather="nccl")  # NCCL all_gather
if rank == 0:
    from oracle import binding as oracle
    od, op, _ = oracle.all_pairs(genes, 3, 2)
    assert res[0] == od, (res[0], od)
    assert np.array_equal(res[1], op)
    assert res2[0] == od and np.array_equal(res2[1], op)
    print("MULTI_OK", res[0][:16])
dist.destroy_process_group()
'''


def test_two_ranks_nccl_digest_matches_oracle(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(SCRIPT.format(root=ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_OK" in out.stdout, out.stdout[-3000:]
