"""Multi-GPU path on the CPU: world_size-2 gloo.  Each rank owns a band of
tile rows (quadray_engine_b200.band_rows, the split qr_render uses across
GPUs), renders it -- here with the oracle standing in for the device -- and
rank 0 gathers the bands; the assembled frame must equal the single-rank one.
This is the host logic of SURVEY.md 8e (tile-row bands + one gather)."""
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
    y0, y1 = pkg.band_rows(h, 8, rank, world)
    band, _, _ = ge.oracle_render(blob, packet=1, y0=y0, y1=y1)
    mine = torch.from_numpy(band[y0:y1].astype(np.int64))
    if rank == 0:
        frame = torch.zeros((h, w), dtype=torch.int64)
        frame[y0:y1] = mine
        for r in range(1, world):
            a, b = pkg.band_rows(h, 8, r, world)
            buf = torch.zeros((b - a, w), dtype=torch.int64)
            dist.recv(buf, src=r)
            frame[a:b] = buf
        np.save(out_path, frame.numpy().astype(np.uint32))
    else:
        dist.send(mine, dst=0)
    dist.barrier()
    dist.destroy_process_group()


def test_band_rows_cover_the_frame(pkg):
    for y_res in (480, 1080, 250, 7, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = [pkg.band_rows(y_res, 8, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == y_res
            for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
                assert a1 == b0 and a0 % 8 == 0 and b0 % 8 == 0


def test_two_ranks_gather_equals_single(tmp_path, entry):
    out = str(tmp_path / "frame.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    _, ref, _ = entry.load_golden("test05_odd")
    assert np.array_equal(got, ref)
