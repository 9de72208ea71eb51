"""Multi-GPU partitioning logic on CPU: world_size-2 gloo processes (the step path itself has
no collective; what is exercised is the member partition and the gather of scan summaries)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from inovesa_b200.ensemble import gather_summaries, member_block, scan_currents  # noqa: E402


def test_member_blocks_partition_the_ensemble():
    for members in (1, 7, 8, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            blocks = [member_block(members, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == members
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        member_block(8, 2, 2)
    c = scan_currents(50e-6, 3e-3, 256)
    assert c.shape == (256,) and c[0] == 50e-6 and abs(c[-1] - 3e-3) < 1e-15


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, members, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle
        from inovesa_b200.params import kara_params
        lo, hi = member_block(members, world, rank)
        currents = scan_currents(20e-6, 400e-6, members)
        oracle = Oracle(False)
        rows = []
        for m in range(lo, hi):  # this rank's members: a tiny current scan on the CPU oracle
            sim = oracle.sim(kara_params(grid_size=32, bunch_current=float(currents[m])))
            sim.step(5)
            g = sim.get()[0]
            rows.append([m, float(currents[m]), float(np.asarray(g, np.float64).sum())])
            sim.close()
        full = gather_summaries(np.array(rows).reshape(hi - lo, 3), members, dist)
        q.put((rank, full))
    finally:
        dist.destroy_process_group()


def test_two_rank_scan_gathers_in_member_order():
    import multiprocessing as mp
    from conftest import _build_oracle
    _build_oracle()
    members, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, members, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # both ranks hold the same table, rows in member order, currents of the scan
    assert np.array_equal(got[0], got[1])
    assert np.array_equal(got[0][:, 0], np.arange(members))
    assert np.allclose(got[0][:, 1], scan_currents(20e-6, 400e-6, members))
    # members differ (different wake strength) and every sum is a finite charge
    assert len(set(np.round(got[0][:, 2], 12))) == members and np.all(np.isfinite(got[0][:, 2]))
