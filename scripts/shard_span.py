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
if os.environ.get("SHARD_FRACS"):  # custom chunk sizes in percent of the tasks, e.g. "12,12,12,8,5,3"
    fr = np.array([float(v) for v in os.environ["SHARD_FRACS"].split(",")])
    bnd = [0] + [int(round(T * v)) for v in np.cumsum(fr) / fr.sum()]
    bnd[-1] = T
    nchunks = len(fr)
    parts = [pkg.partition_all(lens, nranks, bnd[c], bnd[c + 1]) for c in range(nchunks)]
else:
    parts = pkg.partition_chunks(lens, nranks, nchunks)
ids = np.concatenate([parts[c][0] for c in range(nchunks)])
ends = np.cumsum([len(parts[c][0]) for c in range(nchunks)])
cells = sum(lens[i] * lens[j] for i, j in (pkg.task_ij(int(t)) for t in ids))
job = pkg.Job(genes, pxy, pgap, ids)
job.set_chunks(ends)
for it in range(runs):
    t0 = time.perf_counter()
    job.start_async()
    done = []
    for c in range(nchunks):
        job.wait_chunk(c)
        done.append(time.perf_counter() - t0)
    job.finish_async()
    dt = time.perf_counter() - t0
    # what the host fold of an N-GPU job would do with these completion times (0.6 us per task of
    # the whole chunk, in order): time the fold ends after the GPUs' last chunk
    fold_end, prev = 0.0, 0
    for c in range(nchunks):
        fold_end = max(fold_end, done[c]) + 0.6e-6 * nranks * (int(ends[c]) - prev)
        prev = int(ends[c])
    st = job.stats()
    print(f"overlap={os.environ.get('SEQALK_OVERLAP', 'default')} {name} rank 0 of {nranks}, {nchunks} chunks, {len(ids)} tasks: "
          f"span {1e3 * dt:.1f} ms = {cells / dt / 1e9:.0f} GCUPS  (kernel events: fill {st.fill_ms:.1f} tb {st.traceback_ms:.1f} "
          f"dg {st.digest_ms:.1f}, batches {st.batches})  fold model: exposed {1e3 * (fold_end - done[-1]):.1f} ms; "
          f"chunks done at {' '.join(f'{1e3 * d:.0f}' for d in done)}", flush=True)
job.close()
