"""One-process-per-GPU plumbing (torchrun): torch.distributed carries the rendezvous, the barrier and the
ncclUniqueId broadcast; the data path (one all-gather of the row slabs per frame) is NCCL called by
librt_b200 itself on its own stream.  The image shards by interleaved rows: rank g renders
rows y with y % world == g, all sub-samples of a pixel stay on one rank, the scene is replicated.
"""
import os

import numpy as np


def rows_of_rank(H, rank, world):
    """Global row numbers rendered by `rank` (ascending)."""
    return np.arange(rank, H, world)


def rows_per_rank(H, world):
    """Slab height every rank contributes to the all-gather (H padded up to a multiple of world)."""
    return (H + world - 1) // world


def deinterleave(gathered, H, world):
    """gathered: [world, rows_per_rank, W, C] rank-major slabs, as an all-gather leaves them.
    Returns the [H, W, C] image: row y comes from rank y % world, local row y // world."""
    g = np.asarray(gathered)
    assert g.shape[0] == world and g.shape[1] == rows_per_rank(H, world)
    y = np.arange(H)
    return g[y % world, y // world]


def slab_of_rank(image_rows, H, rank, world):
    """Pad this rank's rendered rows ([n_rows, W, C]) to the common slab height."""
    r = rows_per_rank(H, world)
    out = np.zeros((r,) + image_rows.shape[1:], image_rows.dtype)
    out[: len(image_rows)] = image_rows
    return out


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from `src` over the default torch.distributed group (any backend)."""
    import torch
    import torch.distributed as td
    dev = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
    n = torch.tensor([len(payload) if td.get_rank() == src else 0], dtype=torch.int64, device=dev)
    td.broadcast(n, src)
    buf = torch.zeros(int(n.item()), dtype=torch.uint8, device=dev)
    if td.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    td.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def make_renderer():
    """Renderer for this process: single GPU when not under torchrun, otherwise rank `RANK` of `WORLD_SIZE`
    on device LOCAL_RANK with a communicator built from a broadcast ncclUniqueId."""
    from . import binding
    rank, world, local = env_rank_world()
    if world == 1:
        return binding.Renderer(1), 0, 1
    import torch
    import torch.distributed as td
    torch.cuda.set_device(local)
    if not td.is_initialized():
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = binding.Renderer.nccl_unique_id() if rank == 0 else b""
    uid = broadcast_bytes(uid, 0)
    return binding.Renderer(device=local, rank=rank, world=world, nccl_id=uid), rank, world
