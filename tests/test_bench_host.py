"""Host-side logic of bench.py that needs no GPU: the `config` both arms emit, the fixed CPU lattice, and the parity
block (comparison of a frame with the reference's pinned frame) -- single process and a world_size-2 gloo job where
every rank only knows the ids of its own rows."""
import os
import socket
import sys
import zlib

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)


def _fake_pin(path, H=12, W=5, spp=2, seed=3):
    rng = np.random.default_rng(seed)
    ids = rng.integers(-1, 40, (H, W * spp)).astype(np.int32)
    u8 = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    np.savez(path, rows=np.arange(H, dtype=np.int32), id_crc=np.array([zlib.crc32(r.astype("<i4").tobytes()) for r in ids], np.uint32),
             ids_z=np.frombuffer(zlib.compress(ids.astype("<i4").tobytes(), 9), np.uint8), u8=u8, W=W, H=H, pf=1, max_lvl=3)
    return ids, u8


class _FakeRenderer:
    """Stands in for binding.Renderer: returns a prepared frame; a rank of a multi-process job only has its own rows' ids."""
    def __init__(self, ids, u8, rank=0, world=1):
        self.ids, self.u8, self.rank, self.world = ids, u8, rank, world

    def render(self, prm):
        pass

    def download(self, want_prim_id=False):
        ids = self.ids.copy()
        if self.world > 1:
            other = np.arange(len(ids)) % self.world != self.rank
            ids[other] = -2
        return np.zeros(self.u8.shape, np.float32), ids.reshape(-1)

    def download_u8(self):
        return self.u8


def test_config_is_static_and_shared():
    import bench
    from raytracert_b200 import scenes
    s = scenes.balls_standin(grid=8, slices=8, stacks=4)
    a = bench.config_for("balls", "d", s, 800, 800, 4, 3, 1, 8)
    b = bench.config_for("balls", "d", s, 800, 800, 4, 3, 1, 8)
    assert a == b and a["rays_per_pixel"] == 16 and a["triangles"] == s.n_triangles
    assert bench.lattice_pitch(800, 800) == 20 and bench.lattice_pixels(800, 800, 20) == 1600


def test_parity_block_counts_exactly(tmp_path, monkeypatch):
    import bench
    ids, u8 = _fake_pin(tmp_path / "fake.npz")
    monkeypatch.setattr(bench, "pin_path", lambda name: str(tmp_path / (name + ".npz")))
    blk = bench.parity_block(_FakeRenderer(ids, u8), "fake", None, 0, 1, False)
    assert (blk["rows_checked"], blk["id_rows_mismatching"], blk["id_mismatches"], blk["u8_off_by_more_than_1"]) == (12, 0, 0, 0)
    bad_ids, bad_u8 = ids.copy(), u8.astype(int)
    bad_ids[3, 1] += 1; bad_ids[3, 4] += 2; bad_ids[9, 0] = -1 if bad_ids[9, 0] != -1 else 5
    bad_u8[2, 1, 0] += 2 if bad_u8[2, 1, 0] < 200 else -2
    bad_u8[5, 0, 2] += 1 if bad_u8[5, 0, 2] < 200 else -1
    blk = bench.parity_block(_FakeRenderer(bad_ids, bad_u8.astype(np.uint8)), "fake", None, 0, 1, False)
    assert (blk["id_rows_mismatching"], blk["id_mismatches"], blk["u8_off_by_more_than_1"]) == (2, 3, 1)
    assert bench.parity_block(_FakeRenderer(ids, u8), "no_such_workload", None, 0, 1, False)["against"] is None


def _worker(rank, world, port_no, tmp, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as td
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = np.load(os.path.join(tmp, "fake.npz"))
        ids = np.frombuffer(zlib.decompress(z["ids_z"].tobytes()), "<i4").reshape(len(z["rows"]), -1).copy()
        ids[4, 2] += 1      # row 4 belongs to rank 0, row 7 to rank 1: one wrong sample each
        ids[7, 0] += 1
        bench.pin_path = lambda name: os.path.join(tmp, name + ".npz")
        blk = bench.parity_block(_FakeRenderer(ids, z["u8"], rank, world), "fake", None, rank, world, False)
        if rank == 0:
            np.save(out, np.array([blk["rows_checked"], blk["id_rows_mismatching"], blk["id_mismatches"], blk["u8_off_by_more_than_1"]]))
        else:
            assert blk is None
        td.barrier()
    finally:
        td.destroy_process_group()


def test_parity_block_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    _fake_pin(tmp_path / "fake.npz")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    out = str(tmp_path / "blk.npy")
    mp.spawn(_worker, args=(2, port_no, str(tmp_path), out), nprocs=2, join=True)
    assert np.load(out).tolist() == [12, 2, 2, 0]
