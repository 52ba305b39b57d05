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
