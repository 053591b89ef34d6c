"""The N > 1 host logic on CPU: world_size 2 over gloo.  Each rank drives the
host-simulated kernels (tests/hostsim) through the same sharding / gather code
(secp256k1-voi_b200/parallel.py) that the GPU ranks use over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hostsim as hs
        from oracle import oracle as orc
        pkg = importlib.import_module("secp256k1-voi_b200")
        par, synth = pkg.parallel, pkg.synth

        class Be:
            msm_partial = staticmethod(hs.msm_partial)
            msm_combine = staticmethod(hs.msm_combine)
            ecdsa_verify = staticmethod(hs.ecdsa_verify)

        # MSM: global batch of 70 (uneven split: 35/35 at W=2; also checks shard_range)
        n = 70
        w = synth.msm_batch(n, lambda k: orc.batch_scalar_base_mult(k))
        k, p = par.shard_rows(rank, world, w["k32"], w["pt65"])
        out, st = par.msm_sharded(Be, k, p)
        exp, est = orc.scalar_base_mult(w["closed_form_scalar"])
        ok_msm = (st == est) and out.tobytes() == exp
        # an invalid point on ONE rank poisons the result on EVERY rank
        p2 = p.copy()
        if rank == 1:
            p2[0, 64] ^= 1
        out2, st2 = par.msm_sharded(Be, k, p2)
        ok_poison = st2 == 0 and not out2.any()
        # verification: contiguous slices, no collective; concatenation equals the global answer
        v = synth.ecdsa_batch(48, lambda kk: orc.batch_scalar_base_mult(kk))
        mine = par.verify_sharded(Be, rank, world, v["pk65"], v["digest32"], v["sig64"])
        lo, hi = par.shard_range(48, rank, world)
        ok_ver = np.array_equal(mine, v["expected"][lo:hi])
        q.put((rank, bool(ok_msm), bool(ok_poison), bool(ok_ver)))
    except Exception as e:  # report instead of leaving the parent to time out on an empty queue
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    import importlib
    sys.path.insert(0, ROOT)
    par = importlib.import_module("secp256k1-voi_b200").parallel
    for n in (0, 1, 7, 8, 9, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            cuts = [par.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    assert all(len(r) == 4 for r in res), res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_msm, ok_poison, ok_ver in res:
        assert ok_msm and ok_poison and ok_ver, (rank, ok_msm, ok_poison, ok_ver)
