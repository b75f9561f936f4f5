"""GPU span of ONE rank's shard of a sharded job, on one GPU: emulates what a GPU of an N-GPU box
does (chunks of task ids, all chunks enqueued back to back), to compare engine settings without
an N-GPU box:   [SEQALK_OVERLAP=0|1] python scripts/shard_span.py cfg4 8 [nchunks] [runs]"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200")
workloads = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200.workloads")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
nranks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T0 = None
pxy, pgap, genes = workloads.make(name)
lens = [len(g) for g in genes]
T = len(genes) * (len(genes) - 1) // 2
nchunks = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, min(14, T // 16384))
runs = int(sys.argv[4]) if len(sys.argv) > 4 else 3
parts = pkg.partition_chunks(lens, nranks, nchunks)
ids = np.concatenate([parts[c][0] for c in range(nchunks)])
ends = np.cumsum([len(parts[c][0]) for c in range(nchunks)])
cells = sum(lens[i] * lens[j] for i, j in (pkg.task_ij(int(t)) for t in ids))
job = pkg.Job(genes, pxy, pgap, ids)
job.set_chunks(ends)
for it in range(runs):
    t0 = time.perf_counter()
    job.start_async()
    for c in range(nchunks):
        job.wait_chunk(c)
    job.finish_async()
    dt = time.perf_counter() - t0
    st = job.stats()
    print(f"overlap={os.environ.get('SEQALK_OVERLAP', 'default')} {name} rank 0 of {nranks}, {nchunks} chunks, {len(ids)} tasks: "
          f"span {1e3 * dt:.1f} ms = {cells / dt / 1e9:.0f} GCUPS  (kernel events: fill {st.fill_ms:.1f} tb {st.traceback_ms:.1f} "
          f"dg {st.digest_ms:.1f}, batches {st.batches})", flush=True)
job.close()
