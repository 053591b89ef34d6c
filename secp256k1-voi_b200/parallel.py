"""Multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

Verification, base-mult, ScalarMult/ECDH, recovery: items are independent, so
rank g simply owns the contiguous slice [g*n/W, (g+1)*n/W) -- NO collective on
the data path.  Only the MSM has an exchange step: every
rank reduces its slice to one projective partial sum (96 bytes) and a single
all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests) brings the W
partials together; every rank then folds them with W - 1 complete additions
and one inversion.  The message is W * 96 B, i.e. pure latency.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous slice of n items owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(rank, world, *arrays):
    n = len(arrays[0])
    lo, hi = shard_range(n, rank, world)
    return tuple(a[lo:hi] for a in arrays)


def gather_bytes(local_row, group=None, device=None):
    """All-gathers one fixed-size uint8 row per rank -> (world, len) numpy array."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.array(local_row, dtype=np.uint8, copy=True))
    if device is not None:
        t = t.to(device)
    t = t.reshape(-1)
    out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)  # flat: the layout gloo accepts too
    dist.all_gather_into_tensor(out, t, group=group)   # one collective, one copy back
    return out.cpu().numpy().reshape(world, -1)


def init_comm(engine, group=None):
    """Builds the engine's own NCCL communicator (s256_comm_init) over the ranks of `group`: rank 0 draws the
    ncclUniqueId, torch.distributed carries its 128 bytes to the others (the one thing the application has to do),
    every rank joins.  Afterwards engine.msm_sharded runs the gather INSIDE the C ABI."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(engine.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    engine.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)


def msm_sharded(engine, k32_local, pt65_local, vartime=True, group=None, device=None):
    """Point.MultiScalarMult over a batch sharded across ranks: each rank passes ITS slice.
    Returns (out65, status) on every rank.  Invalid points anywhere poison the result.
    An engine with a communicator (init_comm) does everything behind the C ABI (s256_msm_sharded: the partials stay on
    the device, one NCCL all-gather).  Without one -- the gloo tests on CPU, where NCCL does not exist -- the partials
    travel through torch.distributed as 97-byte rows and are folded with msm_combine."""
    if getattr(engine, "comm_size", 0) >= 1:
        return engine.msm_sharded(k32_local, pt65_local, vartime=vartime)
    part, st = engine.msm_partial(k32_local, pt65_local, vartime=vartime)
    row = np.concatenate([np.asarray(part, np.uint8).reshape(96), np.array([st], np.uint8)])
    rows = gather_bytes(row, group=group, device=device)
    if (rows[:, 96] != 1).any():
        return np.zeros(65, np.uint8), 0
    return engine.msm_combine(rows[:, :96])


def verify_sharded(engine, rank, world, pk65, digest32, sig64, flags=0):
    """secec Verify over this rank's contiguous slice of a global batch (no collective)."""
    pk, dg, sg = shard_rows(rank, world, pk65, digest32, sig64)
    return engine.ecdsa_verify(pk, dg, sg, flags)
