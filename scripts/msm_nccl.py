"""Config 5 on N GPUs (run under torchrun): MSM n = 2^LOG2N sharded by contiguous slices,
per-GPU Pippenger partial sums, ONE NCCL all-gather of 97 bytes per rank, combine."""
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
for key in ("k32", "pt65"):  # pinned host buffers, as in bench.py's e2e leg
    w[key] = torch.from_numpy(np.ascontiguousarray(w[key])).pin_memory().numpy()
# the closed form needs the global sum of s_i * d_i: all-reduce it as bytes via gather
mine = int.from_bytes(w["closed_form_scalar"], "big")
if world > 1:
    rows = pkg.parallel.gather_bytes(np.frombuffer(w["closed_form_scalar"], np.uint8), device="cuda")
    total = sum(int.from_bytes(r.tobytes(), "big") for r in rows) % pkg.synth.N
else:
    total = mine
exp, est = eng.scalar_base_mult(np.frombuffer(total.to_bytes(32, "big"), np.uint8))

def run():
    if world > 1:
        return pkg.parallel.msm_sharded(eng, w["k32"], w["pt65"], device="cuda")
    return eng.msm(w["k32"], w["pt65"])

for _ in range(2):
    out, st = run()
assert st == est[0] and np.array_equal(np.asarray(out), exp[0]), "MSM result differs from the closed form"
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    out, st = run()
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"metric": "msm_points_per_sec", "n": n, "n_gpus": world, "value": n * reps / dt.item(),
                      "ms_per_msm": dt.item() / reps * 1e3, "bit_exact_vs_closed_form": True,
                      "note": "pinned host buffers in, 65-byte point out; includes H2D of the slice, the gather and the combine"}))
eng.close()
if world > 1:
    dist.destroy_process_group()
