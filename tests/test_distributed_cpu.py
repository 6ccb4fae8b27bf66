"""world_size-2 gloo test of the multi-process chunk scheduler (one rank per GPU in production; here two
CPU processes with fake engines): contiguous sharding, ordered host-side gather, no data-path collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_windows, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_host_logic import FakeEngine
    from turbo_whisper_workspace_b200.scheduler import DistributedWindowScheduler
    clips = [np.full(100 + i, i / 1000.0, dtype=np.float32) for i in range(n_windows)]
    eng = FakeEngine(f"rank{rank}", max_batch=3)
    sch = DistributedWindowScheduler(eng, rank, world)
    rows = sch.run(clips)
    # max-over-ranks timing reduction used by bench.py (host tensor, gloo)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    torch.save({"rows": rows, "local": sum(eng.calls), "tmax": float(t)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [7, 2, 1])
def test_two_rank_sharding_and_gather(tmp_path, n_windows):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_windows, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    want = [[(100 + i) % 1000, i] for i in range(n_windows)]
    assert r0["rows"] == want and r1["rows"] == want
    assert r0["local"] + r1["local"] == n_windows and abs(r0["local"] - r1["local"]) <= 1
    assert r0["tmax"] == 2.0 and r1["tmax"] == 2.0
