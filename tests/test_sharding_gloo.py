"""Multi-GPU path on the CPU: world_size-2 gloo.  Each rank owns the tile rows
rank, rank + world, ... (quadray_engine_b200.rank_tile_rows, the deal
qr_render_rows(rank, world) renders and qr_render uses across GPUs), renders
them -- here with the oracle standing in for the device -- and rank 0 gathers
them; the assembled frame must equal the single-rank one.  This is the host
logic of SURVEY.md 8e (tile rows dealt round-robin + one gather)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = ge.load_package()
    blob, ref, meta = ge.load_golden("test05_odd")
    h, w = ref.shape
    tile_h = 8
    tls_col = (h + tile_h - 1) // tile_h
    rows_max = (tls_col + world - 1) // world
    # this rank's tile rows, compacted: slot k holds tile row rank + k * world
    mine = torch.zeros((rows_max, tile_h, w), dtype=torch.int64)
    for k, tr in enumerate(pkg.rank_tile_rows(h, tile_h, rank, world)):
        y0, y1 = pkg.tile_row_span(h, tile_h, tr)
        part, _, _ = ge.oracle_render(blob, packet=1, y0=y0, y1=y1)
        mine[k, : y1 - y0] = torch.from_numpy(part[y0:y1].astype(np.int64))
    bufs = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gather_list=bufs, dst=0)
    if rank == 0:
        frame = np.zeros((h, w), dtype=np.uint32)
        for r in range(world):
            for k, tr in enumerate(pkg.rank_tile_rows(h, tile_h, r, world)):
                y0, y1 = pkg.tile_row_span(h, tile_h, tr)
                frame[y0:y1] = bufs[r][k, : y1 - y0].numpy().astype(np.uint32)
        np.save(out_path, frame)
    dist.barrier()
    dist.destroy_process_group()


def test_tile_rows_cover_the_frame(pkg):
    for y_res in (480, 1080, 250, 7, 2160):
        for world in (1, 2, 3, 4, 8):
            tls_col = (y_res + 7) // 8
            seen = sorted(tr for r in range(world) for tr in pkg.rank_tile_rows(y_res, 8, r, world))
            assert seen == list(range(tls_col))
            spans = [pkg.tile_row_span(y_res, 8, tr) for tr in seen]
            assert spans[0][0] == 0 and spans[-1][1] == y_res
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [len(pkg.rank_tile_rows(y_res, 8, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gather_equals_single(tmp_path, entry):
    out = str(tmp_path / "frame.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    _, ref, _ = entry.load_golden("test05_odd")
    assert np.array_equal(got, ref)


def _ring_worker(rank, world, port, shm_path, out_path):
    """The end-to-end protocol of bench.py at N > 1 on the CPU: every rank
    "renders" (oracle) its tile rows of frame k straight into the shared host
    frame k & 1, flags it, rank 0 consumes complete frames and hands the
    buffers back; two frames in flight, several frames of different scenes."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = ge.load_package()
    names = ["test05_odd", "test05_odd_nt", "test05_odd"]       # same geometry, two scene blobs
    blobs = [ge.load_golden(n) for n in names]
    h, w = blobs[0][1].shape
    ring = None
    if rank == 0:
        ring = pkg.SharedFrameRing(shm_path, h, w, world, rank, create=True)
    dist.barrier()
    if rank != 0:
        ring = pkg.SharedFrameRing(shm_path, h, w, world, rank, create=False)
    dist.barrier()
    got = []
    n = 5
    for i in range(n + 1):
        k = i + 1
        if i < n:
            blob = blobs[i % len(blobs)][0]
            ring.wait_free(k)
            if rank == 0 and k > 2:
                assert int(ring.flags[world]) >= k - 2
            for tr in pkg.rank_tile_rows(h, 8, rank, world):
                y0, y1 = pkg.tile_row_span(h, 8, tr)
                part, _, _ = ge.oracle_render(blob, packet=1, y0=y0, y1=y1)
                ring.frame(k)[y0:y1, :w] = part[y0:y1]
        if i > 0:
            ring.mark_done(k - 1)
            if rank == 0:
                ring.wait_complete(k - 1)
                got.append(ring.frame(k - 1).copy())
                ring.release(k - 1)
    if rank == 0:
        np.save(out_path, np.stack(got))
    dist.barrier()
    ring.close(unlink=(rank == 0))
    dist.destroy_process_group()


def test_shared_frame_ring_two_ranks(tmp_path, entry):
    out = str(tmp_path / "frames.npy")
    shm = "/dev/shm/qr_b200_test_%d" % os.getpid()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_ring_worker, args=(2, port, shm, out), nprocs=2, join=True)
    got = np.load(out)
    assert got.shape[0] == 5
    refs = [entry.load_golden(n)[1] for n in ("test05_odd", "test05_odd_nt", "test05_odd")]
    for i in range(5):
        assert np.array_equal(got[i], refs[i % 3]), i
    assert not os.path.exists(shm)
