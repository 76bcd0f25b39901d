"""Row-interleaved sharding, host side, on CPU: world_size-2 gloo processes each render their
rows with the oracle, all-gather rank-major slabs and de-interleave -- the assembled frame must equal the
single-process frame bit for bit.  Also covers the byte broadcast used for the ncclUniqueId."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, bits, load_case, load_scene


def test_deinterleave_roundtrip():
    from raytracert_b200 import dist
    for H, world in [(7, 2), (8, 4), (5, 8), (800, 8), (1, 1), (3, 4)]:
        img = np.arange(H * 3 * 2, dtype=np.float32).reshape(H, 3, 2)
        slabs = np.stack([dist.slab_of_rank(img[dist.rows_of_rank(H, r, world)], H, r, world) for r in range(world)])
        assert slabs.shape[1] == dist.rows_per_rank(H, world)
        assert np.array_equal(dist.deinterleave(slabs, H, world), img)
        got = np.sort(np.concatenate([dist.rows_of_rank(H, r, world) for r in range(world)]))
        assert np.array_equal(got, np.arange(H))


def _worker(rank, world, port_no, case, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as td
    from oracle import pyoracle
    from raytracert_b200 import dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        payload = dist.broadcast_bytes(bytes(range(128)) if rank == 0 else b"", 0)
        assert payload == bytes(range(128))
        c = load_case(case)
        P = pyoracle.PortOracle()
        P.set_scene(load_scene(c["scene"]))
        P.configure(c["eye"], c["lights"], c["features"], c["max_lvl"])
        rgb, _, _ = P.render(c["corners"], c["W"], c["H"], c["pfx"], c["pfy"], y0=rank, ystep=world, threads=2)
        mine = dist.slab_of_rank(rgb[dist.rows_of_rank(c["H"], rank, world)], c["H"], rank, world)
        slabs = [torch.zeros(mine.shape) for _ in range(world)]
        td.all_gather(slabs, torch.from_numpy(mine))
        full = dist.deinterleave(np.stack([s.numpy() for s in slabs]), c["H"], world)
        np.save(out.format(rank=rank), full)
        td.barrier()
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("case", ["shadow_test_2lights_lvl3", "dodge_48x27"])
def test_two_ranks_gloo(built, tmp_path, case):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    out = str(tmp_path / "frame_{rank}.npy")
    mp.spawn(_worker, args=(2, port_no, case, out), nprocs=2, join=True)
    c = load_case(case)
    for r in range(2):
        assert np.array_equal(bits(np.load(out.format(rank=r))), bits(c["rgb"]))
