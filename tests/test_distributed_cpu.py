"""CPU, world_size 2, gloo: trajectory-parallel sharding + final gather (the only communication)."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _fake_generate(x0, u0, scenarios, T, traj_id0):
    """stand-in for ClosedLoopGenerator.generate: deterministic function of the GLOBAL trajectory id only."""
    B = len(x0)
    ids = np.arange(traj_id0, traj_id0 + B, dtype=np.float64)
    clean = ids[:, None, None] + np.arange(T + 1)[None, :, None] * 0.01 + np.arange(6)[None, None, :] * 1e-3 + x0[:, None, :]
    return {"clean": clean, "noisy": clean + 0.5, "U": np.zeros((B, T, 2)) + ids[:, None, None],
            "iters_total": ids.astype(np.int64) * T}


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from trajectory_generation_b200 import distributed as tgd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, T = 7, 4
    x0 = np.arange(B * 6, dtype=np.float64).reshape(B, 6)
    u0 = np.zeros((B, 2))
    lo, hi, local = tgd.generate_sharded(_fake_generate, x0, u0, [None] * B, T, rank, world)
    out = tgd.gather_results(local, lo, hi, B, dist)

    class FakeOpenLoop:          # stand-in for OpenLoopGenerator: same generate(x0, T, traj_id0) contract
        def generate(self, x0_, T_, traj_id0=0):
            return _fake_generate(x0_, None, None, T_, traj_id0)
    lo2, hi2, local2 = tgd.generate_openloop_sharded(FakeOpenLoop(), x0, T, rank, world)
    out2 = tgd.gather_results(local2, lo2, hi2, B, dist)
    if rank == 0:
        full = _fake_generate(x0, u0, None, T, 0)
        ret["ok"] = all(np.array_equal(out[k], full[k]) for k in full) and all(np.array_equal(out2[k], full[k]) for k in full)
        ret["shapes"] = {k: out[k].shape for k in out}
    else:
        ret["other"] = out is None and out2 is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_equals_single_rank():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret["ok"] and ret["other"]
        assert ret["shapes"]["clean"] == (7, 5, 6)
