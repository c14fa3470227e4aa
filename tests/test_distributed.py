"""Host-side multi-GPU logic on CPU: tile partition + the two collectives, world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_tile_partition_covers_image_once():
    from bihrt import multi
    for (w, h) in ((200, 120), (1920, 1080), (33, 65), (32, 32)):
        for world in (1, 2, 3, 4, 8):
            own = multi.tile_owner(w, h, world)
            assert own.shape == (h, w) and own.min() >= 0 and own.max() < world
            total = sum(multi.shard_ray_count(w, h, 4, r, world) for r in range(world))
            assert total == w * h * 4
            # a tile is never split between ranks
            assert np.all(own[:32, :min(32, w)] == own[0, 0])
    # round-robin balance: 1080p over 8 ranks within one tile of even
    cnt = [multi.shard_ray_count(1920, 1080, 1, r, 8) for r in range(8)]
    assert max(cnt) - min(cnt) <= 32 * 32 * 2


class FakeRenderer:
    """Stands in for bihrt.Renderer in the collectives test: the 'BIH blob' is a byte pattern."""

    def __init__(self, rank):
        self.rank, self.device = rank, 0
        self.blob = np.arange(1000, dtype=np.uint8) if rank == 0 else None
        self.imported = None

    def bih_blob_bytes(self):
        return len(self.blob)

    def bih_export(self, buf, nbytes):
        import torch
        buf.copy_(torch.from_numpy(self.blob))

    def bih_import(self, buf, nbytes):
        self.imported = buf.numpy().copy()

    def sync(self):
        pass

    # stream ordering with torch's current stream (events in the real Renderer): record the calls
    def wait_torch(self):
        self.__dict__.setdefault("order_log", []).append("ctx_waits_torch")

    def torch_wait(self):
        self.__dict__.setdefault("order_log", []).append("torch_waits_ctx")

    # in-place replication: the blob region of an n-triangle scene is the broadcast buffer itself
    def bih_region_tensor(self, n):
        import torch
        if getattr(self, "region", None) is None:
            self.region = torch.from_numpy((np.arange(64 + 64 * n) % 251).astype(np.uint8)) if self.rank == 0 else torch.zeros(64 + 64 * n, dtype=torch.uint8)
        return self.region

    def bih_adopt(self, n):
        self.adopted = n

    # fused gather set-up: rank 0 exports a 64-byte handle, the others open it
    ipc_fails = False

    def framebuffer_ipc_export(self, w, h):
        if self.ipc_fails:
            raise RuntimeError("no IPC here")
        return bytes((7 * i + w + h) % 256 for i in range(64))

    def framebuffer_ptr(self):
        return 0xF00D, 0, 0

    def framebuffer_ipc_open(self, handle):
        self.opened = bytes(handle)
        return 0xBEEF


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from bihrt import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r = FakeRenderer(rank)
        n = multi.replicate_bih(r, dist, src=0, device="cpu")
        ok_blob = n == 1000 and (rank == 0 or np.array_equal(r.imported, np.arange(1000, dtype=np.uint8)))
        nb = multi.replicate_bih_inplace(r, dist, 37, src=0)
        ok_blob = ok_blob and nb == 64 + 64 * 37 and np.array_equal(r.region.numpy(), (np.arange(nb) % 251).astype(np.uint8)) \
            and (rank == 0 or getattr(r, "adopted", None) == 37) and (rank != 0 or not hasattr(r, "adopted"))
        # the in-place broadcast is bracketed by the two stream orderings (torch waits for the build, the context waits for the collective)
        ok_blob = ok_blob and r.order_log[-2:] == ["torch_waits_ctx", "ctx_waits_torch"]
        tok = torch.ones(1, dtype=torch.int32)
        multi.frame_barrier(dist, tok)
        ok_blob = ok_blob and int(tok.item()) == world
        # fused-gather set-up: every rank ends up with a pointer (own / peer), or all agree there is none
        ptr, is_peer = multi.open_peer_framebuffer(r, dist, 40, 30, dst=0, device="cpu")
        ok_blob = ok_blob and ((rank == 0 and (ptr, is_peer) == (0xF00D, False)) or
                               (rank == 1 and (ptr, is_peer) == (0xBEEF, True) and r.opened == bytes((7 * i + 70) % 256 for i in range(64))))
        r.ipc_fails = True
        ptr, is_peer = multi.open_peer_framebuffer(r, dist, 40, 30, dst=0, device="cpu")
        ok_blob = ok_blob and ptr is None and not is_peer
        # framebuffer gather: each rank holds its own tiles of a known image, 0 elsewhere
        w, h = 200, 120
        full = (np.arange(w * h, dtype=np.int32).reshape(h, w) % 9973) + 1
        mine = np.where(multi.tile_owner(w, h, world) == rank, full, 0).astype(np.int32)
        t = torch.from_numpy(mine.copy())
        multi.gather_framebuffer(t, dist, dst=0)
        ok_fb = rank != 0 or np.array_equal(t.numpy(), full)
        q.put((rank, bool(ok_blob), bool(ok_fb)))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True, True), (1, True, True)]
