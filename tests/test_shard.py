"""Host-side sharding logic, including the N > 1 path on CPU (gloo, world_size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raymarchdenoisercuda_b200 import shard


def test_assign_sequences_covers_every_sequence_once():
    for n, w in [(8, 1), (8, 2), (8, 4), (8, 8), (5, 3), (0, 2)]:
        parts = shard.assign_sequences(n, w)
        assert len(parts) == w
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_row_bands_partition_and_halo():
    for H, w in [(4320, 2), (4320, 4), (4320, 8), (1080, 7), (17, 17)]:
        bands = shard.row_bands(H, w)
        assert bands[0].row0 == 0 and bands[-1].row0 + bands[-1].rows == H
        for a, b in zip(bands, bands[1:]):
            assert a.row0 + a.rows == b.row0
        assert bands[0].halo_top == 0 and bands[-1].halo_bot == 0
        assert max(b.rows for b in bands) - min(b.rows for b in bands) <= 1
    assert shard.frame_halo(5) == 2 + 4 + 8 + 16 + 32 + 1 + 3
    assert shard.frame_halo(0) == 3
    b = shard.row_bands(4320, 8)[3]
    assert b.halo_top == 66 and b.halo_bot == 66
    with pytest.raises(ValueError):
        shard.row_bands(3, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.assign_sequences(8, world)[rank]
        seconds = 1.0 + rank  # the slow rank decides
        mx = shard.max_over_ranks(seconds)
        agg = shard.aggregate_mpixels_per_s(len(mine) * 1e6, seconds)
        band = shard.row_bands(4320, world)[rank]
        rows = shard.sum_over_ranks(band.rows)
        out[rank] = (mx, agg, rows)
    finally:
        dist.destroy_process_group()


def test_two_rank_reduction_gloo():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        mx, agg, rows = out[r]
        assert mx == 2.0                      # max over ranks
        assert abs(agg - 8e6 / 2.0 / 1e6) < 1e-9  # all pixels over the slowest rank's time
        assert rows == 4320


def _swap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a 48-row "frame" of 8 columns whose value encodes (row, frame): every rank holds its band + a 4-row halo
        H, W, halo = 48, 8, 4
        band = shard.row_bands(H, world)[rank]
        top, bot = min(halo, band.row0), min(halo, H - band.row0 - band.rows)
        ok = True
        for frame in range(3):
            full = (torch.arange(H, dtype=torch.float32)[:, None] * 100 + frame).repeat(1, W)
            own = full[band.row0:band.row0 + band.rows]
            send_up = own[:halo].contiguous() if top else None      # my first rows are the upper neighbour's bottom halo
            send_dn = own[-halo:].contiguous() if bot else None     # my last rows are the lower neighbour's top halo
            recv_up = torch.full((halo, W), -1.0) if top else None
            recv_dn = torch.full((halo, W), -1.0) if bot else None
            shard.neighbour_swap(rank, send_up, recv_up, send_dn, recv_dn)
            if top:
                ok &= bool(torch.equal(recv_up, full[band.row0 - halo:band.row0]))
            if bot:
                ok &= bool(torch.equal(recv_dn, full[band.row0 + band.rows:band.row0 + band.rows + halo]))
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_halo_swap_gloo(world):
    """The exchange step of the row-band scheme on CPU: after the swap every rank holds exactly its neighbours'
    boundary rows (first and last band have one neighbour, a middle band two), for several frames in a row."""
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_swap_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
