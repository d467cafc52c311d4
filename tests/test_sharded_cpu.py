"""Host-side logic of the multi-GPU paths on CPU: geometry, and the two exchange steps of the row-band
path over gloo with world_size 2 and 3 (the same torch.distributed code runs over NCCL on GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from canny_edge_b200 import sharded


def test_frame_slice_covers_everything():
    for n, world in [(512, 8), (10, 4), (3, 8), (7, 1)]:
        got = [sharded.frame_slice(n, r, world) for r in range(world)]
        assert got[0][0] == 0 and sum(c for _, c in got) == n
        for (a, ca), (b, _) in zip(got, got[1:]):
            assert a + ca == b


def test_band_geometry():
    g = [sharded.band_geometry(32768, 32768, r, 8, 1.4) for r in range(8)]
    assert all(x.rows == 4096 and x.halo == 7 for x in g)           # 8 bands of 4096 rows, 7-row halo
    assert g[0].halo_above == 0 and g[0].halo_below == 7 and g[7].halo_below == 0 and g[3].buffer_rows == 4096 + 14
    assert [x.row0 for x in g] == [4096 * r for r in range(8)]
    u = [sharded.band_geometry(1001, 64, r, 3, 5.0) for r in range(3)]   # uneven split, 17-row halo
    assert [x.rows for x in u] == [334, 334, 333] and u[1].halo == 17 and u[2].row0 == 668
    with pytest.raises(ValueError):
        sharded.band_geometry(40, 64, 0, 8, 1.4)                       # bands shorter than the halo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, sigma, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img = torch.from_numpy(np.random.default_rng(5).integers(0, 256, (H, W)).astype(np.uint8))
        geo = sharded.band_geometry(H, W, rank, world, sigma)
        band = img[geo.row0:geo.row0 + geo.rows].clone()
        buf = sharded.exchange_halos(band, geo)
        want = img[geo.row0 - geo.halo_above:geo.row0 + geo.rows + geo.halo_below]
        ok_halo = bool((buf == want).all()) and buf.shape[0] == geo.buffer_rows
        rec = torch.full((24,), rank + 1, dtype=torch.uint8)
        allrec = sharded.gather_records(rec, world)
        ok_rec = allrec.tolist() == [r + 1 for r in range(world) for _ in range(24)]
        out_q.put((rank, ok_halo, ok_rec))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,sigma", [(2, 64, 1.4), (3, 100, 1.4), (2, 90, 5.0)])
def test_band_exchanges_over_gloo(world, H, sigma):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, 48, sigma, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _, _ in res) == list(range(world))
    assert all(h and g for _, h, g in res), res
