"""Kernel times of one job under several knob settings, in ONE process on one box:
   python scripts/ab_env.py cfg5 3 "" "SEQALK_TB_BAND=0,SEQALK_HOST_HASH=0" ...
Every setting gets a fresh Job (a job reads the knobs when it is planned); the result checksum must agree."""
import hashlib, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200")
workloads = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200.workloads")
name, runs, settings = sys.argv[1], int(sys.argv[2]), sys.argv[3:] or [""]
pxy, pgap, genes = workloads.make(name)
pkg.init(1)
sums = set()
for setting in settings:
    env = dict(kv.split("=") for kv in setting.split(",") if kv)
    for k, v in env.items():
        os.environ[k] = v
    job = pkg.Job(genes, pxy, pgap)
    best = None
    for it in range(runs):
        t0 = time.perf_counter(); job.run(); t1 = time.perf_counter()
        st = job.stats()
        row = (1e3 * (t1 - t0), st.fill_ms, st.traceback_ms, st.digest_ms)
        best = row if best is None or row[0] < best[0] else best
    pen, dig = job.fetch()
    cs = hashlib.sha256(pen.tobytes() + dig.tobytes()).hexdigest()[:16]
    sums.add(cs)
    print(f"{name} [{setting or 'default'}] best of {runs}: wall {best[0]:.2f} ms  fill {best[1]:.2f}  traceback {best[2]:.2f}  "
          f"digest {best[3]:.2f}  launches tb {st.traceback_launches}  checksum {cs}", flush=True)
    job.close()
    for k in env:
        os.environ.pop(k, None)
print("checksums agree" if len(sums) == 1 else f"CHECKSUMS DIFFER: {sums}")
