"""Times the phases of the one-process-per-GPU end-to-end call (run under torchrun)."""
import importlib, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "multiple-sequence-alignment-openmp-openmpi_b200"
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dmod = importlib.import_module(PKG + ".distributed"); workloads = importlib.import_module(PKG + ".workloads")
pxy, pgap, genes = workloads.make(sys.argv[1] if len(sys.argv) > 1 else "cfg4")
lens = [len(g) for g in genes]
for it in range(3):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter(); plan = dmod.ShardPlan(lens, world, 8)
    t1 = time.perf_counter(); shard = dmod.GpuShard(genes, pxy, pgap, plan, rank, local)
    t2 = time.perf_counter(); shard.start(); res = dmod.run_sharded(plan, rank, shard.run_chunk); shard.finish()
    torch.cuda.synchronize(); t3 = time.perf_counter(); shard.close(); t4 = time.perf_counter()
    if rank == 0:
        st = None
        print(f"it{it} plan {1e3*(t1-t0):.1f} ms, shard(create job: pack+H2D+plan) {1e3*(t2-t1):.1f} ms, run {1e3*(t3-t2):.1f} ms, close {1e3*(t4-t3):.1f} ms, total {1e3*(t4-t0):.1f} ms", flush=True)
if world > 1: dist.destroy_process_group()
