"""Row-tile sharding of frames across the GPUs of one box.

Every op on the hot path is per-pixel independent (colorlut, hsvfilter, hsvdetector) or a per-block sum
(blockhash), so a frame splits by contiguous rows with no halo: rank r owns rows
[row_range(H, N, r)).  No collective is needed when each rank returns its tile to the host over its own
PCIe link; a device-side consumer that wants the whole frame does ONE all-gather of the tiles.
The block-hash variant aligns tiles to hash-block rows and all-reduces the 64 partial sums.
"""
from __future__ import annotations


def row_range(height: int, world: int, rank: int, align: int = 1):
    """rows [r0, r1) owned by `rank`: ceil(H / N) rows rounded up to `align`, last tile may be short/empty"""
    per = -(-height // world)
    per = -(-per // align) * align
    r0 = min(height, rank * per)
    return r0, min(height, r0 + per)


def tile_view(frame, world: int, rank: int, align: int = 1):
    """the row tile of a (H, stride) array / tensor"""
    r0, r1 = row_range(frame.shape[0], world, rank, align)
    return frame[r0:r1]


def all_gather_rows(dist, tile, height: int, world: int, align: int = 1):
    """reassemble a (H, stride) uint8 frame on every rank from the ranks' row tiles (torch tensors).
    Tiles are padded to the common tile height so a single all_gather_into_tensor / all_gather suffices."""
    import torch
    per = row_range(height, world, 0, align)[1]
    stride = tile.shape[1]
    padded = tile
    if tile.shape[0] != per:
        padded = torch.zeros((per, stride), dtype=tile.dtype, device=tile.device)
        padded[: tile.shape[0]] = tile
    padded = padded.contiguous()
    out = torch.empty((world * per, stride), dtype=tile.dtype, device=tile.device)
    if tile.device.type == "cuda":
        dist.all_gather_into_tensor(out.view(-1), padded.view(-1))
    else:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded)
        out = torch.cat(parts, dim=0)
    return out[:height]


def all_reduce_sums(dist, sums):
    """blockhash: partial u32 block sums of the row tiles -> full-frame sums (exact: integer adds)"""
    import torch
    t = sums.to(torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class DeviceBuffer:
    """a raw device allocation seen through __cuda_array_interface__ (torch.as_tensor(buf, device=...) wraps it, no copy)"""

    def __init__(self, ptr: int, shape, owner=None):
        self.ptr, self.shape, self._owner = ptr, tuple(shape), owner
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerFrames:
    """Frame buffers for the fused colorlut + all-gather kernel (b200vfx_colorlut_process_tile_gather).

    Every rank owns `nbuf` whole-frame buffers (height x stride bytes) and one flag block in peer-mappable memory; the
    CUDA IPC handles travel through `dist.all_gather_object` (plumbing only -- the data path is the kernel's own
    NVLink stores).  process(tile) writes this rank's rows into buffer k = epoch % nbuf of EVERY rank; when the call
    completes on the context stream, frame(k) on this rank holds all ranks' rows."""

    def __init__(self, ctx, dist, height: int, stride: int, nbuf: int = 2, align: int = 1, multicast: bool = False):
        self.ctx, self.dist, self.height, self.stride, self.nbuf, self.align = ctx, dist, height, stride, nbuf, align
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.epoch = 0
        self.mc = [0] * nbuf         # multicast address of buffer k (0 = unicast stores to every peer)
        self._own_flags = ctx.peer_alloc(256)
        if multicast and self.world > 1:
            self._init_multicast()
            return
        self._own = [ctx.peer_alloc(height * stride) for _ in range(nbuf)]
        mine = {"frames": [h for _, h in self._own], "flags": self._own_flags[1]}
        everyone = [mine]
        if self.world > 1:
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine)
        self._opened = []
        self.frames = [[0] * self.world for _ in range(nbuf)]   # [buffer][rank] -> address usable on this device
        self.flags = [0] * self.world
        for r, info in enumerate(everyone):
            if r == self.rank:
                for k in range(nbuf):
                    self.frames[k][r] = self._own[k][0]
                self.flags[r] = self._own_flags[0]
                continue
            for k in range(nbuf):
                p = ctx.peer_open(info["frames"][k]); self._opened.append(p); self.frames[k][r] = p
            p = ctx.peer_open(info["flags"]); self._opened.append(p); self.flags[r] = p
        if self.world > 1:
            dist.barrier()

    def _init_multicast(self):
        """frame buffers from torch's symmetric memory (plumbing: cuMemCreate + cuMulticastBindMem + the handle exchange
        between the processes); the flag blocks stay ordinary peer memory.  Raises when the fabric has no multicast."""
        import torch
        import torch.distributed._symmetric_memory as symm
        dist, nbuf = self.dist, self.nbuf
        self._symm, self._own, self._opened = [], [], []
        self.frames = [[0] * self.world for _ in range(nbuf)]
        for k in range(nbuf):
            dev = torch.cuda.current_device()
            t = symm.empty(self.height * self.stride, dtype=torch.uint8, device=torch.device("cuda", dev))
            h = symm.rendezvous(t, dist.group.WORLD)
            if not h.multicast_ptr:
                raise RuntimeError("no NVSwitch multicast on this box")
            self._symm.append((t, h))
            self._own.append((t.data_ptr(), None))
            self.frames[k] = [int(p) for p in h.buffer_ptrs]
            self.mc[k] = int(h.multicast_ptr)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, self._own_flags[1])
        self.flags = [0] * self.world
        for r, hdl in enumerate(everyone):
            if r == self.rank:
                self.flags[r] = self._own_flags[0]
            else:
                p = self.ctx.peer_open(hdl); self._opened.append(p); self.flags[r] = p
        dist.barrier()

    def rows(self):
        return row_range(self.height, self.world, self.rank, self.align)

    def process(self, width: int, tile, tile_stride: int):
        """colorlut on this rank's row tile (device tensor) -> rows of buffer (epoch % nbuf) on every rank; returns k"""
        r0, r1 = self.rows()
        self.epoch += 1
        k = self.epoch % self.nbuf
        self.ctx.colorlut_process_tile_gather("RGBA", width, r1 - r0, tile, tile_stride, self.world, self.rank,
                                              self.frames[k], self.stride, r0, self.flags, self.epoch, multicast=self.mc[k])
        return k

    def frame(self, k: int):
        """this rank's whole-frame buffer k as a __cuda_array_interface__ object of shape (height, stride)"""
        return DeviceBuffer(self._own[k][0], (self.height, self.stride), owner=self)

    def status(self) -> int:
        """synchronise and return the epoch of the last timed-out peer wait (0 = none)"""
        return self.ctx.peer_status(self._own_flags[0])

    def close(self):
        if self.world > 1:
            self.ctx.synchronize()
            self.dist.barrier()
        for p in self._opened:
            self.ctx.peer_close(p)
        self._opened = []
        if self.world > 1:
            self.dist.barrier()
        for p, h in self._own:
            if h is not None:        # symmetric-memory buffers belong to torch
                self.ctx.peer_free(p)
        self._symm = []
        self.ctx.peer_free(self._own_flags[0])
        self._own = []
