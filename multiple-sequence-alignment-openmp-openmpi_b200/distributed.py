"""One process per GPU: the reference's master/worker task farm re-shaped for NVLink.

Reference (submit/xuliny-seqalkway.cpp): rank 0 broadcasts the inputs (:249-266), hands out
Triple{i,j,id} tasks one by one (:280-331), workers return Packet{penalty,id,hash} (:342-357,
:399-411), rank 0 folds the hashes in id order (:334-337).

Here: every rank holds the inputs (replicated, as in the reference), the k(k-1)/2 tasks are
sharded statically by cell count (LPT — independent units, no data-path collective during the
fill), every rank runs its shard on its own GPU through the C ABI, and the per-task
(penalty, 64-byte digest) records are gathered to rank 0 once per chunk of task ids.  Rank 0
scatters them back into the reference's task order and folds the chain; with ``nchunks > 1`` the
fold of chunk c overlaps the GPUs' work on chunk c+1.

Two gathers (34 MB per k = 1000 job, so neither is a bandwidth question):
* ``gather="host"`` (default on GPUs): each rank copies its chunk's records to pinned host memory
  with the copy engine and the ranks meet in a gloo ``gather`` — no kernel is launched for the
  hand-off.  The fill is a persistent kernel that owns every SM; a NCCL kernel queued behind it
  would not get a CTA slot before the fill's tail (with the engine's overlapped batch pipelines
  there is always a fill resident), which delayed every chunk by one fill in the measurements.
* ``gather="nccl"``: ``all_gather_into_tensor`` over NVLink on the GPUs, kept for comparison.
The single-process form (``seqalk_init(N)`` + ``seqalk_all_pairs``) gathers with peer stores from
the digest kernel instead (csrc/api.cu).

``compute`` is injectable so that the shard/gather/reorder logic is testable on CPU with the
gloo backend (tests/test_distributed_cpu.py); the default computes on the rank's GPU and
fails loudly without one.
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import Job, chunk_bounds, fold_chain, partition_chunks

ComputeFn = Callable[[Sequence[bytes], int, int, np.ndarray], Tuple[np.ndarray, np.ndarray]]


class ShardPlan:
    """Chunks of task ids x ranks.  chunk c covers ids [bounds[c], bounds[c+1])."""

    def __init__(self, lens: Sequence[int], world: int, nchunks: int = 1):
        k = len(lens)
        T = k * (k - 1) // 2
        nchunks = max(1, min(nchunks, max(T, 1)))
        self.T, self.world = T, world
        self.bounds = chunk_bounds(T, nchunks)  # tapered tail: the last fold is the exposed one
        self.ids: List[List[np.ndarray]] = partition_chunks(lens, world, nchunks)  # one parallel LPT pass

    @property
    def nchunks(self) -> int:
        return len(self.ids)

    def rank_ids(self, rank: int) -> np.ndarray:
        """All ids of a rank, chunk after chunk (each chunk increasing)."""
        parts = [self.ids[c][rank] for c in range(self.nchunks)]
        return np.concatenate(parts) if parts else np.zeros(0, np.int64)

    def max_count(self, c: int) -> int:
        return max((len(a) for a in self.ids[c]), default=0)


class GpuShard:
    """A rank's tasks resident on its GPU: ONE job (one set of device buffers) cut into the
    plan's chunks; results land in two torch tensors covering all of the rank's tasks."""

    def __init__(self, genes: Sequence[bytes], pxy: int, pgap: int, plan: ShardPlan, rank: int, device: int):
        import torch

        self.torch = torch
        self.device = torch.device("cuda", device)
        self.plan, self.rank = plan, rank
        counts = [len(plan.ids[c][rank]) for c in range(plan.nchunks)]
        self.ends = np.cumsum(counts).astype(np.int64)
        self.starts = self.ends - np.asarray(counts, dtype=np.int64)
        ids = plan.rank_ids(rank)
        self.n = len(ids)
        self.job = Job(genes, pxy, pgap, ids, device=device) if self.n else None
        if self.job is not None:
            self.job.set_chunks(self.ends)
        self.pen = torch.zeros(max(self.n, 1), dtype=torch.int32, device=self.device)
        self.dig = torch.zeros((max(self.n, 1), 64), dtype=torch.uint8, device=self.device)
        # gather buffers: [penalty int32 | 64 digest bytes] = 68 bytes per task, padded to the
        # largest shard of the chunk so that one all_gather_into_tensor moves everything
        self.send = [torch.zeros((max(plan.max_count(c), 1), 68), dtype=torch.uint8, device=self.device)
                     for c in range(plan.nchunks)]
        # host path: pinned mirrors of the result arrays + a copy stream (copy engine only)
        self.h_pen = torch.zeros(max(self.n, 1), dtype=torch.int32).pin_memory()
        self.h_dig = torch.zeros((max(self.n, 1), 64), dtype=torch.uint8).pin_memory()
        self.h_send = [torch.zeros((max(plan.max_count(c), 1), 68), dtype=torch.uint8) for c in range(plan.nchunks)]
        self.copy_stream = torch.cuda.Stream(device=self.device)

    def start(self):
        """Start every chunk (asynchronously, back to back on the GPU)."""
        if self.job is not None:
            self.job.start_async(self.pen.data_ptr(), self.dig.data_ptr())

    def run_chunk(self, c: int):
        """Waits (on torch's current stream, not on the host) for this rank's results of chunk
        c and packs them into the chunk's send buffer."""
        lo, hi = int(self.starts[c]), int(self.ends[c])
        n = hi - lo
        if n:
            self.job.wait_chunk(c, self.torch.cuda.current_stream(self.device).cuda_stream)
            self.send[c][:n, :4] = self.pen[lo:hi].view(self.torch.uint8).view(n, 4)
            self.send[c][:n, 4:] = self.dig[lo:hi]
        return self.send[c]

    def run_chunk_host(self, c: int):
        """Host form of run_chunk: waits for chunk c, copies its records device -> pinned host with
        the copy engine (no kernel launch) and packs them into a CPU tensor (maxcount_c, 68)."""
        lo, hi = int(self.starts[c]), int(self.ends[c])
        n = hi - lo
        buf = self.h_send[c]
        if n:
            torch = self.torch
            self.job.wait_chunk(c, self.copy_stream.cuda_stream)
            with torch.cuda.stream(self.copy_stream):
                self.h_pen[lo:hi].copy_(self.pen[lo:hi], non_blocking=True)
                self.h_dig[lo:hi].copy_(self.dig[lo:hi], non_blocking=True)
            self.copy_stream.synchronize()
            out = buf.numpy()
            out[:n, :4] = self.h_pen[lo:hi].numpy().view(np.uint8).reshape(n, 4)
            out[:n, 4:] = self.h_dig[lo:hi].numpy()
        return buf

    def finish(self):
        if self.job is not None:
            self.job.finish_async()

    def stats(self):
        return [self.job.stats()] if self.job is not None else []

    def close(self):
        if self.job is not None:
            self.job.close()
            self.job = None


def _scatter(plan: ShardPlan, c: int, gathered: np.ndarray, penalties: np.ndarray, digests: np.ndarray) -> None:
    """gathered: (world, maxcount, 68) uint8 -> task-ordered arrays."""
    for r in range(plan.world):
        ids = plan.ids[c][r]
        n = len(ids)
        if n:
            rec = gathered[r, :n]
            penalties[ids] = rec[:, :4].copy().view(np.int32).reshape(n)
            digests[ids] = rec[:, 4:]


_HOST_GROUPS = {}


def host_group(group=None):
    """A gloo group over the same ranks as ``group`` (created once; collective call)."""
    import torch.distributed as dist

    key = id(group) if group is not None else 0
    if key not in _HOST_GROUPS:
        if dist.get_backend(group) == "gloo":
            _HOST_GROUPS[key] = group
        else:
            ranks = dist.get_process_group_ranks(group) if group is not None else None
            _HOST_GROUPS[key] = dist.new_group(ranks=ranks, backend="gloo")
    return _HOST_GROUPS[key]


def run_sharded(plan: ShardPlan, rank: int, run_chunk: Callable[[int], "object"], group=None,
                ) -> Optional[Tuple[str, np.ndarray]]:
    """Drives the chunks: run, gather, and (rank 0) scatter + fold in task order.

    ``run_chunk(c)`` returns a torch uint8 tensor (maxcount_c, 68): on the device the process
    group communicates on (NCCL / gloo ``all_gather_into_tensor``), or on the CPU while the group
    is a GPU one (host gather: gloo ``gather`` to rank 0).  Returns (digest, penalties) on rank 0,
    None elsewhere."""
    import torch
    import torch.distributed as dist

    world = plan.world
    penalties = np.zeros(plan.T, dtype=np.int32) if rank == 0 else None
    digests = np.zeros((plan.T, 64), dtype=np.uint8) if rank == 0 else None
    state = {"hex": "", "error": None}
    # rank 0: one folder thread takes the gathered chunks in order, so that the loop below never
    # waits for a fold (the chunks of a GPU complete in pairs, see chunk_bounds in csrc/api.cu)
    todo: "queue.Queue" = queue.Queue()

    def fold_loop():
        while True:
            item = todo.get()
            if item is None:
                return
            c, host = item
            try:
                _scatter(plan, c, host, penalties, digests)
                b, e = plan.bounds[c], plan.bounds[c + 1]
                state["hex"] = fold_chain(digests[b:e], state["hex"])
            except Exception as exc:  # surfaced after the loop
                state["error"] = exc
                return

    folder: Optional[threading.Thread] = None
    if rank == 0:
        folder = threading.Thread(target=fold_loop)
        folder.start()

    try:
        for c in range(plan.nchunks):
            send = run_chunk(c)
            host = None
            if world > 1 and send.device.type == "cpu" and dist.get_backend(group) != "gloo":
                # host gather: the records are already in host memory on every rank
                hg = host_group(group)
                parts = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
                dist.gather(send, parts, dst=dist.get_global_rank(hg, 0), group=hg)
                if rank == 0:
                    host = np.stack([p.numpy() for p in parts])
            elif world > 1:
                rows = send.shape[0]
                flat = torch.empty((world * rows, send.shape[1]), dtype=send.dtype, device=send.device)
                dist.all_gather_into_tensor(flat, send.contiguous(), group=group)
                recv = flat.view(world, rows, send.shape[1])
                if rank == 0:
                    host = recv.cpu().numpy()  # D2H of the gathered records (synchronises the stream)
            elif rank == 0:
                host = send.unsqueeze(0).cpu().numpy()
            if rank == 0:
                todo.put((c, host))
    finally:
        if folder is not None:
            todo.put(None)
            folder.join()
    if rank == 0:
        if state["error"] is not None:
            raise state["error"]
        return state["hex"], penalties
    return None


def all_pairs_distributed(genes: Sequence[bytes], pxy: int, pgap: int, *, nchunks: int = 1, group=None,
                          device: Optional[int] = None, compute: Optional[ComputeFn] = None,
                          gather: str = "host") -> Optional[Tuple[str, np.ndarray]]:
    """getMinimumPenalties over all ranks of the default process group.

    Every rank passes the same inputs.  Returns (digest, penalties) on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lens = [len(g) for g in genes]
    plan = ShardPlan(lens, world, nchunks)
    if compute is None:
        if not torch.cuda.is_available():
            raise RuntimeError("all_pairs_distributed needs one B200 per rank (no CPU fallback)")
        dev = torch.cuda.current_device() if device is None else device
        shard = GpuShard(genes, pxy, pgap, plan, rank, dev)
        try:
            shard.start()
            res = run_sharded(plan, rank, shard.run_chunk_host if gather == "host" else shard.run_chunk, group)
            shard.finish()
            return res
        finally:
            shard.close()

    def run_chunk(c: int):
        ids = plan.ids[c][rank]
        buf = np.zeros((max(plan.max_count(c), 1), 68), dtype=np.uint8)
        if len(ids):
            pen, dig = compute(genes, pxy, pgap, ids)
            buf[:len(ids), :4] = np.asarray(pen, dtype=np.int32).view(np.uint8).reshape(-1, 4)
            buf[:len(ids), 4:] = np.asarray(dig, dtype=np.uint8).reshape(-1, 64)
        return torch.from_numpy(buf)

    return run_sharded(plan, rank, run_chunk, group)
