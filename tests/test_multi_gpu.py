"""The sharded MSM with its NCCL all-gather INSIDE the C ABI (s256_msm_sharded), two ranks on two GPUs.
Skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu` runs it."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, log2n, q):
    sys.path.insert(0, ROOT)
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import oracle as orc
        pkg = importlib.import_module("secp256k1-voi_b200")
        n = 1 << log2n
        lo, hi = pkg.parallel.shard_range(n, rank, world)
        eng = pkg.Engine(device=rank, max_batch=max(hi - lo, 1024))
        pkg.parallel.init_comm(eng)
        w = pkg.synth.msm_batch(hi - lo, eng.scalar_base_mult, start=lo)
        rows = pkg.parallel.gather_bytes(np.frombuffer(w["closed_form_scalar"], np.uint8), device="cuda")
        total = sum(int.from_bytes(r.tobytes(), "big") for r in rows) % pkg.synth.N
        exp, est = orc.scalar_base_mult(total.to_bytes(32, "big"))
        out, st = eng.msm_sharded(w["k32"], w["pt65"])                        # host pointers, one sync
        ok_host = (st == est) and out.tobytes() == exp
        dk, dp = torch.from_numpy(w["k32"]).cuda(), torch.from_numpy(w["pt65"]).cuda()
        out, st = eng.msm_sharded(dk, dp)                                     # device resident, no sync
        ok_dev = int(st.cpu()[0]) == est and out.cpu().numpy().tobytes() == exp
        bad = w["pt65"].copy()
        if rank == world - 1:
            bad[3, 64] ^= 1                                                   # one undecodable point on ONE rank
        out, st = eng.msm_sharded(w["k32"], bad)
        ok_poison = st == 0 and not out.any()
        empty = eng.msm_sharded(w["k32"][:0] if rank else w["k32"], w["pt65"][:0] if rank else w["pt65"])   # ragged: rank > 0 has nothing
        q.put((rank, bool(ok_host), bool(ok_dev), bool(ok_poison), int(empty[1])))
        eng.close()
    except Exception as e:
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("log2n", [12, 18])
def test_msm_sharded_two_gpus_nccl_inside_the_c_abi(log2n):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, log2n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert all(len(r) == 5 for r in res), res
    for rank, ok_host, ok_dev, ok_poison, st_empty in res:
        assert ok_host and ok_dev and ok_poison and st_empty == 1, res
