"""Multi-GPU rendering: one process per GPU (torchrun), `torch.distributed` over NCCL for the one exchange step.

The reference distributes 48x48 buckets to threads through an atomic cursor (/root/reference/src/main.cpp:323-371,
src/sdl.cpp:243-262) and every thread writes its own pixels of the shared `vfb`. Across GPUs the same decomposition is
static and needs exactly one collective at the end of the frame:

* ``tiles``   -- rank r owns the 8x4 pixel tiles t of the frame with t % world == r, renders all samples of them and
                 leaves zeros elsewhere; the partial frames are summed onto rank 0 (``reduce(SUM)``; disjoint pixels, so
                 the sum is exact) and divided by spp there.
* ``samples`` -- rank r renders samples [r*spp/world, (r+1)*spp/world) of EVERY pixel (possible because the counter-based
                 RNG makes sample i of pixel p independent of who renders it); ``reduce(SUM)`` then / spp. Best load
                 balance; the result equals the single-GPU frame up to FP32 summation order.

* ``p2p``     -- the tile split without the reduction: rank 0 exports its frame (CUDA IPC), every rank maps it and its
                 kernels store the finished pixels of their own tiles straight into it over NVLink; a one-element
                 all-reduce on the render streams orders "everyone has written" before "rank 0 reads". No partial frames, no
                 clearing, no 1.9 MB reduce, no resolve pass. The exported allocation holds TWO frames used alternately: the
                 peers' stores of frame k+1 go to the other one while rank 0 may still be copying frame k to the host, and they
                 cannot start frame k+2 before the all-reduce of frame k+1, which rank 0 enters only after that copy (stream
                 order). Falls back to ``tiles`` where peer mapping is unavailable.

* ``host``    -- the tile split when the frame is wanted in HOST memory (``render()``): the frame lives in a POSIX
                 shared-memory mapping that every rank process maps and page-locks (fray_gpu_host_register), and every rank's
                 GPU stores the finished pixels of its own tiles straight into it over its own PCIe link
                 (fray_gpu_render_to_host). No partial frames, no reduce, no 1.9 MB device-to-host copy through one link; the
                 only exchange is a pair of counters per rank in the same mapping ("share k is in memory" / "rank 0 has
                 entered frame k"). Two frames are used alternately, as in ``p2p``. Bit-identical to the single-GPU frame
                 whenever the tile split is. ``render_device()`` (a frame in rank 0's DEVICE memory) runs the ``samples`` or
                 ``tiles`` reduction in this mode.

The scene is replicated (the largest bundled scene is ~25 MB). There is no collective inside the render path.
"""
from __future__ import annotations

import os

import numpy as np

import fray_b200 as fb

CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: the default stream as an explicit handle


def shard(rank: int, world: int, spp: int, mode: str) -> dict:
    """Frame keyword arguments (see FrayGpuFrame) that select rank `rank`'s share of the frame."""
    if world <= 1:
        return {}
    if mode == "tiles":
        return dict(bucket_rank=rank, bucket_count=world)
    if mode == "samples":
        if spp < world:
            raise ValueError(f"cannot split {spp} samples over {world} ranks; use mode='tiles'")
        return dict(sample_begin=(rank * spp) // world, sample_end=((rank + 1) * spp) // world)
    raise ValueError(f"unknown shard mode {mode!r}")


def choose_mode(spp: int, world: int) -> str:
    return "samples" if spp >= 8 * world else "tiles"


def reduce_partials(partial, spp: int, group=None):
    """Sum per-rank partial SUM frames (torch tensors, any device) onto rank 0 and resolve: returns sum / spp on rank 0,
    None elsewhere. With an un-initialised process group this is the single-process identity."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(partial, dst=0, op=dist.ReduceOp.SUM, group=group)
        if dist.get_rank(group) != 0:
            return None
    return partial / float(spp)


class DistributedRenderer:
    """One rank of a multi-GPU render. Rank 0 ends up with the finished frame."""

    def __init__(self, scene: fb.Scene, mode: str = "auto", precision: int = fb.FP32, device: int | None = None):
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.dist = dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
        torch.cuda.set_device(self.device)
        self.scene = scene
        self.spp = scene.spp
        self.mode = choose_mode(self.spp, self.world) if mode == "auto" else mode
        self.ctx = fb.GpuContext(scene, self.device, precision)
        h, w = scene.height, scene.width
        self.shape = (h, w, 3)
        self.peer_frame = 0  # p2p: address of rank 0's pair of frames as seen from this GPU
        self.frames_rendered = 0  # p2p: frame k goes to buffer k % 2
        self.frame = None
        if self.mode == "p2p" and self.world > 1 and not self._setup_p2p():
            self.mode = "tiles"
        if self.mode == "p2p" and self.world == 1:
            self.mode = "tiles"
        self.device_mode = self.mode  # what render_device() does
        self.shm = None
        if self.mode == "host":
            self.device_mode = choose_mode(self.spp, self.world)
            if self.world == 1 or not self._setup_host():
                self.mode = self.device_mode
        if self.mode != "p2p":
            self.partial = torch.zeros(self.shape, dtype=torch.float32, device=f"cuda:{self.device}")
            self.frame = torch.zeros_like(self.partial) if self.rank == 0 else None
        else:
            self.token = torch.zeros(1, dtype=torch.float32, device=f"cuda:{self.device}")
        self.host_frame = torch.empty(self.shape, dtype=torch.float32).pin_memory() if self.rank == 0 else None

    def _setup_p2p(self) -> bool:
        """Rank 0 exports its frame, everyone maps it. Collective: all ranks agree on success or fall back together."""
        torch, dist = self.torch, self.dist
        dev = f"cuda:{self.device}"
        handle = torch.zeros(64, dtype=torch.uint8, device=dev)
        ok = 1
        if self.rank == 0:
            try:
                self.peer_frame, raw = self.ctx.frame_export()
                handle.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
            except fb.FrayError:
                ok = 0
        dist.broadcast(handle, src=0)
        if self.rank != 0:
            try:
                self.peer_frame = self.ctx.frame_import(bytes(handle.cpu().numpy().tobytes()))
            except fb.FrayError:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if self.rank != 0 and self.peer_frame:
                self.ctx.frame_close(self.peer_frame)
            self.peer_frame = 0
            return False
        if self.rank == 0:
            # a tensor view of the exported frame for reads on rank 0 (the memory belongs to the context)
            n = self.shape[0] * self.shape[1] * self.shape[2]
            self.frame_pair = self._wrap_device_pointer(self.peer_frame, 2 * n).view((2,) + self.shape)
            self.frame = self.frame_pair[0]
        return True

    # ---- ``host`` mode: one frame in shared, page-locked host memory ---------------------------------------------------------
    _HOST_HEADER = 4096  # bytes in front of the two frames: per rank one 64-byte line {done, entered}

    def _setup_host(self) -> bool:
        """Rank 0 creates the mapping, everyone maps and page-locks it. Collective: all ranks succeed or fall back together."""
        import mmap
        torch, dist = self.torch, self.dist
        nbytes = self._HOST_HEADER + 2 * self.shape[0] * self.shape[1] * self.shape[2] * 4
        name = [None]
        if self.rank == 0:
            name[0] = f"/dev/shm/fray_frame_{os.getpid()}_{os.environ.get('MASTER_PORT', '0')}"
            with open(name[0], "wb") as f:
                f.truncate(nbytes)
        dist.broadcast_object_list(name, src=0)
        ok = 1
        try:
            fd = os.open(name[0], os.O_RDWR)
            try:
                self.shm = mmap.mmap(fd, nbytes)
            finally:
                os.close(fd)
            self.shm_np = np.frombuffer(self.shm, dtype=np.uint8)
            self.shm_addr = self.shm_np.ctypes.data
            fb.host_register(self.shm_addr, nbytes)
        except (OSError, ValueError, fb.FrayError) as e:
            import sys
            print(f"fray_b200.dist: rank {self.rank}: no shared host frame ({e}); falling back to the {self.device_mode} reduction", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{self.device}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        all_ok = int(flag.item())  # (waits for the collective: every rank is past its open() now)
        if self.rank == 0:
            os.unlink(name[0])  # every rank has it open (or has given up): the name can go
        if all_ok == 0:
            self._close_host()
            return False
        self.shm_flags = self.shm_np[:self._HOST_HEADER].view(np.int64).reshape(-1, 8)  # [rank] = {done, entered, ...}
        n = self.shape[0] * self.shape[1] * self.shape[2]
        self.shm_frames = self.shm_np[self._HOST_HEADER:].view(np.float32).reshape((2,) + self.shape)
        self.shm_frame_bytes = n * 4
        self.host_frames = 0
        return True

    def _close_host(self):
        if self.shm is not None:
            try:
                fb.host_unregister(self.shm_addr)
            except Exception:
                pass
            self.shm_flags = self.shm_frames = self.shm_np = None
            try:
                self.shm.close()
            except BufferError:
                pass  # a caller still holds a view of the last frame
            self.shm = None

    def _render_host(self, seed: int):
        """Frame k goes to buffer k % 2. A rank may start frame k only after rank 0 has ENTERED frame k (rank 0's caller is then
        done with frame k - 2, whose buffer this is); rank 0 returns frame k once every rank has reported its share in memory."""
        import time
        dbg = os.environ.get("FRAY_DIST_DEBUG")
        t0 = time.perf_counter()
        k = self.host_frames = self.host_frames + 1
        flags = self.shm_flags
        if self.rank == 0:
            flags[0, 1] = k
        else:
            while flags[0, 1] < k:
                pass
        which = k & 1
        stream = self.torch.cuda.current_stream().cuda_stream or CUDA_STREAM_LEGACY
        t1 = time.perf_counter()
        self.ctx.render_to_host(self.shm_addr + self._HOST_HEADER + which * self.shm_frame_bytes, stream, spp=self.spp, seed=seed,
                                bucket_rank=self.rank, bucket_count=self.world)
        t2 = time.perf_counter()
        self.torch.cuda.current_stream().synchronize()  # this share is in host memory
        t3 = time.perf_counter()
        flags[self.rank, 0] = k
        if self.rank != 0:
            return None
        for r in range(1, self.world):
            while flags[r, 0] < k:
                pass
        t4 = time.perf_counter()
        if dbg:
            self.dbg = getattr(self, "dbg", [])
            self.dbg.append(((t1 - t0) * 1e6, (t2 - t1) * 1e6, (t3 - t2) * 1e6, (t4 - t3) * 1e6))
        return self.shm_frames[which]

    def _wrap_device_pointer(self, ptr: int, count: int):
        """A float32 torch tensor over `count` floats of device memory owned by the context (no copy)."""
        torch = self.torch

        class _Iface:  # the CUDA array interface is the supported way to alias foreign device memory
            __cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_Iface(), device=f"cuda:{self.device}")

    def render_device(self, seed: int = 42):
        """Kernels + the one exchange step, everything enqueued on torch's current stream. Rank 0: self.frame holds the image."""
        torch = self.torch
        # torch's default stream has handle 0, which the C ABI reads as "the context's own stream": name it explicitly
        stream = torch.cuda.current_stream().cuda_stream or CUDA_STREAM_LEGACY
        mode = self.device_mode if self.mode == "host" else self.mode
        if mode == "p2p":
            which = self.frames_rendered & 1  # double buffering: see the module docstring
            self.frames_rendered += 1
            target = self.peer_frame + which * self.shape[0] * self.shape[1] * self.shape[2] * 4
            self.ctx.render_device(target, stream, spp=self.spp, seed=seed, flags=fb.FRAME_OWNED_ONLY,
                                   bucket_rank=self.rank, bucket_count=self.world)
            self.dist.all_reduce(self.token)  # stream-ordered barrier: every rank's stores precede rank 0's reads
            if self.rank == 0:
                self.frame = self.frame_pair[which]
            return
        kw = shard(self.rank, self.world, self.spp, mode)
        self.ctx.render_device(self.partial.data_ptr(), stream, spp=self.spp, seed=seed, flags=fb.FRAME_SUM, **kw)
        if self.world > 1:
            self.dist.reduce(self.partial, dst=0, op=self.dist.ReduceOp.SUM)
        if self.rank == 0:
            self.ctx.resolve_device(self.partial.data_ptr(), self.frame.data_ptr(), self.spp, stream)

    def render(self, seed: int = 42) -> np.ndarray | None:
        """End to end: render, exchange, and bring the frame to (pinned) host memory on rank 0. With a reduction, rank 0's
        resolve pass stores the quotients straight into the pinned frame (fray_gpu_resolve_to_host): no device-to-host copy."""
        if self.mode == "host":
            return self._render_host(seed)
        if self.mode == "p2p":
            self.render_device(seed)
            if self.rank != 0:
                return None
            self.host_frame.copy_(self.frame, non_blocking=True)
            self.torch.cuda.current_stream().synchronize()
            return self.host_frame.numpy()
        torch = self.torch
        stream = torch.cuda.current_stream().cuda_stream or CUDA_STREAM_LEGACY
        kw = shard(self.rank, self.world, self.spp, self.mode)
        self.ctx.render_device(self.partial.data_ptr(), stream, spp=self.spp, seed=seed, flags=fb.FRAME_SUM, **kw)
        if self.world > 1:
            self.dist.reduce(self.partial, dst=0, op=self.dist.ReduceOp.SUM)
        if self.rank != 0:
            return None
        self.ctx.resolve_to_host(self.partial.data_ptr(), self.host_frame.data_ptr(), self.spp, stream)
        torch.cuda.current_stream().synchronize()
        return self.host_frame.numpy()

    def stats(self) -> fb.RenderStats:
        return self.ctx.sync()

    def close(self):
        if self.mode == "p2p" and self.rank != 0 and self.peer_frame:
            self.torch.cuda.synchronize()
            self.ctx.frame_close(self.peer_frame)
            self.peer_frame = 0
        self.frame = None
        if self.mode == "host":
            self.torch.cuda.synchronize()
            self._close_host()
        self.ctx.close()
