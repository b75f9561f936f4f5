"""Phase times of the end-to-end C-ABI call (seqalk_all_pairs from host buffers) on one GPU:
   python scripts/e2e_profile.py cfg2 [calls]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200")
workloads = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200.workloads")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pxy, pgap, genes = workloads.make(name)
pkg.init(1)
for it in range(calls):
    t0 = time.perf_counter(); d, p = pkg.get_minimum_penalties(genes, pxy, pgap); t1 = time.perf_counter()
    prof = pkg.last_profile()
    print(f"{name} call {it}: python wall {1e3 * (t1 - t0):.2f} ms  " + "  ".join(f"{k} {v:.2f}" for k, v in prof.items() if v), flush=True)
