"""Kernel times of one job, for same-box A/B of two library builds:
   SEQALK_LIB=ab/x/libseqalk.so python scripts/ab_lib.py cfg4 [runs]"""
import hashlib, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200")
workloads = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200.workloads")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pxy, pgap, genes = workloads.make(name)
pkg.init(1)
job = pkg.Job(genes, pxy, pgap)
for it in range(runs):
    t0 = time.perf_counter(); job.run(); t1 = time.perf_counter()
    st = job.stats()
    print(f"{os.environ.get('SEQALK_LIB', 'default')} {name} run{it}: wall {1e3 * (t1 - t0):.1f} ms  fill {st.fill_ms:.1f}  "
          f"traceback {st.traceback_ms:.1f}  digest {st.digest_ms:.1f}  batches {st.batches}", flush=True)
pen, dig = job.fetch()
print("  result checksum", hashlib.sha256(pen.tobytes() + dig.tobytes()).hexdigest()[:16], "chain", pkg.fold_chain(dig)[:16])
job.close()
