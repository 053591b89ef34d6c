"""Where does a sharded MSM call spend its time?  Under torchrun: per rank, the local Pippenger alone (s256_msm_dev on
the slice, no communicator) against s256_msm_sharded_dev (same slice + pack + ncclAllGather + fold + encode)."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module("secp256k1-voi_b200")
n = 1 << int(os.environ.get("LOG2N", "20"))
lo, hi = pkg.parallel.shard_range(n, rank, world)
eng = pkg.Engine(device=local, max_batch=max(hi - lo, 1024))
w = pkg.synth.msm_batch(hi - lo, eng.scalar_base_mult, start=lo)
dk, dp = torch.from_numpy(w["k32"]).cuda(), torch.from_numpy(w["pt65"]).cuda()

def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    t_enq = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, t_enq

res = {"rank": rank, "n_local": hi - lo}
res["local_msm_dev_ms"], res["local_enqueue_ms"] = timed(lambda: eng.msm(dk, dp))
if world > 1:
    pkg.parallel.init_comm(eng)
    res["sharded_dev_ms"], res["sharded_enqueue_ms"] = timed(lambda: eng.msm_sharded(dk, dp))
    t = torch.zeros(112, dtype=torch.uint8, device="cuda"); o = torch.zeros(112 * world, dtype=torch.uint8, device="cuda")
    res["torch_allgather_112B_ms"], _ = timed(lambda: dist.all_gather_into_tensor(o, t))
print(json.dumps(res), flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
