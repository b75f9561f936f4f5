"""In-process multi-GPU job (seqalk_init(N) + seqalk_all_pairs) with different chunk counts:
   python scripts/peer_chunks.py cfg4 8 10,14,18 [runs]      (SEQALK_TRACE=1 for the phase marks)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200")
workloads = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200.workloads")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
ngpus = int(sys.argv[2]) if len(sys.argv) > 2 else 2
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
runs = int(sys.argv[4]) if len(sys.argv) > 4 else 3
pxy, pgap, genes = workloads.make(name)
cells = sum(len(genes[i]) * len(genes[j]) for i in range(len(genes)) for j in range(i))
pkg.init(ngpus)
pkg.get_minimum_penalties(genes, pxy, pgap)  # warm-up: buffers, peer mappings
for nch in variants:
    if nch:
        os.environ["SEQALK_PEER_CHUNKS"] = str(nch)
    else:
        os.environ.pop("SEQALK_PEER_CHUNKS", None)
    for it in range(runs):
        sys.stderr.write(f"[peer_chunks] chunks={nch or 'default'} run {it}\n"); sys.stderr.flush()
        t0 = time.perf_counter()
        digest, pen = pkg.get_minimum_penalties(genes, pxy, pgap)
        ms = 1e3 * (time.perf_counter() - t0)
        pr = pkg.last_profile()
        print(f"{name} x{ngpus} chunks={nch or 'default'} run{it}: call {ms:.1f} ms = {cells / ms / 1e6:.0f} GCUPS  "
              + "  ".join(f"{k}={v:.1f}" for k, v in pr.items()) + f"  digest {digest[:16]}", flush=True)
