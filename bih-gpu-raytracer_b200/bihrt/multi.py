"""Multi-GPU plumbing: one process per GPU, torch.distributed for the two collectives the path has.

The reference is single-GPU (no NCCL/MPI/peer copies).  Rays are independent given a
replicated tree, so the partition is data-parallel:

  * the BIH is built once on rank `src` and replicated with ONE broadcast of the blob
    [header | nodes | leaf-ordered triangles] (bihrt_bih_export / bihrt_bih_import);
  * ray batches are partitioned by UNIT INTERLEAVE (preferred: every rank walks every 32x32 tile and traces every
    G-th 32-ray unit of it -- a few neighbouring pixels with all their samples -- writing per-pixel hit counts,
    bihrt_render_interleaved: balanced by construction, single-GPU locality, samples stay in consecutive lanes), by SAMPLE (rank r traces samples [spp*r/G, spp*(r+1)/G) of every
    pixel and stores per-pixel hit counts, bihrt_render_samples -- every rank walks the whole image, so
    the warps keep the single-GPU coherence; the default when spp >= G) or by TILE (32x32-pixel tiles
    dealt round-robin, tile k -> rank k mod G, other pixels 0, bihrt_render_shard; used when spp < G);
  * the framebuffer is gathered either INSIDE the trace kernel (unit interleave, bihrt_render_interleaved_to: every
    rank stores the final colour of the pixels it owns straight into the gathering GPU's framebuffer, mapped
    through CUDA IPC, so the transfer rides NVLink while the kernel runs and only a barrier follows), or, for
    the sample / tile partitions, with ONE reduce (SUM of counts, or of disjoint shards) followed for sample
    sharding by bihrt_framebuffer_resolve on the destination rank.

No collective runs during traversal.  The helpers take any torch.distributed backend so the same code
is exercised with gloo on CPU tensors in tests/test_distributed.py.
"""
import numpy as np

TILE = 32


def tile_owner(w, h, world):
    """(h, w) int32 array: rank that renders each pixel (tile id = ty * tiles_x + tx, id mod world)."""
    tx = (w + TILE - 1) // TILE
    jj, ii = np.meshgrid(np.arange(h) // TILE, np.arange(w) // TILE, indexing="ij")
    return ((jj * tx + ii) % world).astype(np.int32)


def sample_range(spp, rank, world):
    """Sample sharding: rank -> [begin, end) of the spp samples of every pixel (contiguous, balanced)."""
    return (spp * rank) // world, (spp * (rank + 1)) // world


def shard_ray_count(w, h, spp, rank, world):
    return int((tile_owner(w, h, world) == rank).sum()) * spp


class _CudaView:
    """Zero-copy torch view of a device pointer owned by the library (via __cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def framebuffer_tensor(renderer):
    """The context's framebuffer (Renderer::m_cudaDestResource) as a (h, w) int32 CUDA tensor, no copy."""
    import torch
    ptr, w, h = renderer.framebuffer_ptr()
    return torch.as_tensor(_CudaView(ptr, (h, w), "<i4"), device="cuda:%d" % renderer.device)


def replicate_bih(renderer, dist, src=0, device=None):
    """Broadcast the BIH built on rank `src` to every rank (one collective).  Returns the blob size."""
    import torch
    rank = dist.get_rank()
    dev = device if device is not None else "cuda:%d" % renderer.device
    size = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        size[0] = renderer.bih_blob_bytes()
    dist.broadcast(size, src=src)
    nbytes = int(size.item())
    blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if rank == src:
        renderer.wait_torch()
        renderer.bih_export(blob, nbytes)
        renderer.sync()
    dist.broadcast(blob, src=src)
    if rank != src:
        renderer.wait_torch()            # the import reads the blob the collective wrote on torch's current stream
        renderer.bih_import(blob, nbytes)
        renderer.sync()
    return nbytes


def open_peer_framebuffer(renderer, dist, w, h, dst=0, device=None):
    """Fused gather: rank `dst` shares its w x h framebuffer with the other ranks (CUDA IPC handle, one 65-byte
    broadcast); every rank gets the device pointer to hand to Renderer.render_interleaved_to, whose trace kernel
    then stores the finished pixels straight into rank dst's memory over NVLink.  Returns (pointer, is_peer);
    pointer is None on a rank where the mapping is not available (no exception escapes before or after the
    broadcast, so the ranks stay in step and can agree on a fall-back)."""
    import torch
    rank = dist.get_rank()
    dev = device if device is not None else "cuda:%d" % renderer.device
    h65 = torch.zeros(65, dtype=torch.uint8)
    if rank == dst:
        try:
            h65[:64] = torch.frombuffer(bytearray(renderer.framebuffer_ipc_export(w, h)), dtype=torch.uint8)
            h65[64] = 1
        except Exception:       # noqa
            pass
    h65 = h65.to(dev)
    dist.broadcast(h65, src=dst)
    h65 = h65.cpu()
    if int(h65[64]) != 1:
        return None, False
    if rank == dst:
        return renderer.framebuffer_ptr()[0], False
    try:
        return renderer.framebuffer_ipc_open(bytes(h65[:64].numpy().tobytes())), True
    except Exception:           # noqa
        return None, False


def frame_barrier(dist, token):
    """The fused gather has no data collective: the frame on rank dst is complete when every rank's launch has
    finished.  A one-element all-reduce on the launching stream orders that (token: 1-element CUDA tensor)."""
    dist.all_reduce(token)


def replicate_bih_inplace(renderer, dist, n, src=0):
    """Broadcast the BIH of an n-triangle scene built on rank `src` straight between the contexts' blobs: one
    collective, no staging copies, no host synchronisation (n must be known on every rank).  Returns the bytes."""
    import torch
    if hasattr(renderer, "bih_region_tensor"):          # test doubles hand out their own (CPU) tensor
        blob = renderer.bih_region_tensor(n)
        nbytes = blob.numel()
    else:
        ptr, nbytes = renderer.bih_region(n)
        blob = torch.as_tensor(_CudaView(ptr, (nbytes,), "|u1"), device="cuda:%d" % renderer.device)
    # The collective runs on torch's current stream, the build and the trace on the context's: order them with events
    # (both are no-ops when the context was given torch's current stream, as bench.py does).
    renderer.torch_wait()                # the build that wrote the blob / the trace that still reads the old one
    dist.broadcast(blob, src=src)
    renderer.wait_torch()                # whatever the context does next sees the broadcast blob
    if dist.get_rank() != src:
        renderer.bih_adopt(n)
    return nbytes


def gather_framebuffer(fb, dist, dst=0, renderer=None):
    """Combine the ranks' disjoint shards (0 outside a rank's tiles) on rank `dst`: one reduce.  With `renderer` the
    reduce (torch's current stream) is ordered after the render and before the context's next call."""
    if renderer is not None:
        renderer.torch_wait()
    dist.reduce(fb, dst=dst, op=dist.ReduceOp.SUM)
    if renderer is not None:
        renderer.wait_torch()
    return fb


def frame_barrier_ordered(renderer, dist, token):
    """frame_barrier for a context that does not run on torch's current stream."""
    renderer.torch_wait()
    dist.all_reduce(token)
    renderer.wait_torch()
