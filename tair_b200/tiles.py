"""Tile driver pieces of val_patches.py: split, blend/stitch, and the data-parallel sharding of tiles over GPUs.

``split_image_with_overlap`` / ``merge_patches_with_overlap`` keep the reference signatures and results
(val_patches.py:25-92, :114-206); the merge is ONE kernel (``tair_blend_tiles``) instead of a Python loop of
2 slice-adds per tile, bit-identical in fp32.  New here (the reference restores tiles one by one on one device and
does not shard by rank, val_patches.py:230-231,296,316): ``shard_tiles`` / ``gather_tiles`` distribute the row-major
tile list round-robin over ranks — tiles are independent until the blend, so there is no inter-step communication —
and reassemble the decoded tiles with a single NCCL ``all_gather_into_tensor`` before the blend.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

LQ_PATCH, LQ_OVERLAP = 128, 16  # geometry hard-coded inside the reference merge (val_patches.py:133-148)


def tile_grid(height: int, width: int, patch_size: int = LQ_PATCH, overlap: int = LQ_OVERLAP) -> Tuple[int, int, int, int]:
    """(rows, cols, padded_height, padded_width) of the tile grid covering a height x width image."""
    stride = patch_size - overlap
    rows = math.ceil((height - overlap) / stride)
    cols = math.ceil((width - overlap) / stride)
    return rows, cols, (rows - 1) * stride + patch_size, (cols - 1) * stride + patch_size


def split_image_with_overlap(image, patch_size: int = LQ_PATCH, overlap: int = LQ_OVERLAP):
    """PIL image (or HWC / HW uint8 array) -> list of PIL tiles, left-to-right, top-to-bottom; the image is
    zero-padded on the right / bottom so that every tile is full (val_patches.py:25-92)."""
    from PIL import Image
    arr = np.asarray(image)
    h, w = arr.shape[:2]
    rows, cols, ph, pw = tile_grid(h, w, patch_size, overlap)
    canvas = np.zeros((ph, pw) + arr.shape[2:], dtype=arr.dtype)
    canvas[:h, :w] = arr
    stride = patch_size - overlap
    tiles = []
    for r in range(rows):
        for c in range(cols):
            t = canvas[r * stride:r * stride + patch_size, c * stride:c * stride + patch_size]
            tiles.append(Image.fromarray(t.astype(np.uint8)))
    return tiles


# ---- PIL-exact bicubic resampling tables (host) and the GPU tile front-end ----------------------------------------

def pil_bicubic_coeffs(in_size: int, out_size: int):
    """Per-output-index windows and 22-bit fixed-point coefficients of PIL's Image.resize(BICUBIC) on 8-bit images
    (Pillow src/libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc, bicubic a = -0.5, box = whole
    tile).  Returns (bounds [out,2] int32 = first tap / number of taps, coeffs [out,ksize] int32)."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(np.ceil(support)) * 2 + 1

    def cubic(x):
        a = -0.5
        x = abs(x)
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
        if x < 2.0:
            return (((x - 5.0) * x + 8.0) * x - 4.0) * a
        return 0.0
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / fscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)           # C (int) cast: truncation; operands are positive here
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [cubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(k)
        if ww != 0.0:
            k = [v / ww for v in k]
        bounds[xx] = (xmin, xmax)
        for x, v in enumerate(k):
            coeffs[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
    return bounds, coeffs


class TileFrontEnd:
    """GPU replacement of split_image_with_overlap + per-tile PIL resize + ToTensor (val_patches.py:25-92,291-294,318):
    the zero-padded LQ image is uploaded once; ``tiles(idx)`` crops and resizes any subset of tiles on the device."""

    def __init__(self, image, device, patch_size: int = LQ_PATCH, overlap: int = LQ_OVERLAP, out_size: int = 512):
        arr = np.asarray(image)
        if arr.ndim == 2:
            arr = np.repeat(arr[:, :, None], 3, axis=2)
        h, w = arr.shape[:2]
        self.rows, self.cols, ph, pw = tile_grid(h, w, patch_size, overlap)
        canvas = np.zeros((ph, pw, 3), np.uint8)
        canvas[:h, :w] = arr[:, :, :3]
        self.image = torch.from_numpy(canvas).to(device)
        stride = patch_size - overlap
        self.origins = torch.tensor([[r * stride, c * stride] for r in range(self.rows) for c in range(self.cols)],
                                    dtype=torch.int32, device=device)
        b, c = pil_bicubic_coeffs(patch_size, out_size)
        self.bounds, self.coeffs = torch.from_numpy(b).to(device), torch.from_numpy(c).to(device)
        self.patch_size, self.out_size = patch_size, out_size

    def __len__(self) -> int:
        return self.rows * self.cols

    def tiles(self, indices) -> torch.Tensor:
        idx = torch.as_tensor(list(indices), dtype=torch.long, device=self.image.device)
        return ops.tiles_bicubic(self.image, self.origins.index_select(0, idx), self.bounds, self.coeffs,
                                 self.patch_size, self.out_size)


def merge_patches_with_overlap(patches: Sequence[torch.Tensor], original_size: Tuple[int, int], patch_size: int = 512,
                               overlap: int = 64) -> torch.Tensor:
    """list of (1,3,P,P) CUDA tensors (row-major tile order) -> (1,3,scale*H,scale*W) blended image.
    As in the reference the grid is derived from the LQ geometry 128/16 regardless of ``patch_size``/``overlap``."""
    tiles = patches if isinstance(patches, torch.Tensor) else torch.cat(list(patches), 0)
    h, w = original_size
    rows, cols, _, _ = tile_grid(h, w)
    scale = patch_size / LQ_PATCH
    return ops.blend_tiles(tiles.float().contiguous(), rows, cols, overlap, int(h * scale), int(w * scale))


# ---- data-parallel sharding ------------------------------------------------------------------------------------

def shard_tiles(n_tiles: int, rank: int, world_size: int) -> List[int]:
    """Global tile indices owned by ``rank``: p with p % world_size == rank.  Per-tile noise must be
    keyed by the global index so results do not depend on world_size."""
    return list(range(rank, n_tiles, world_size))


def tiles_per_rank(n_tiles: int, world_size: int) -> int:
    return (n_tiles + world_size - 1) // world_size


def gather_tiles(local: torch.Tensor, n_tiles: int, group=None) -> torch.Tensor:
    """local: this rank's decoded tiles [len(shard_tiles(...)), C, P, P] in shard order.  Returns all ``n_tiles``
    tiles in global row-major order on every rank, using one all_gather_into_tensor (ragged tail zero-padded)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = tiles_per_rank(n_tiles, world)
    pad = per - local.shape[0]
    if pad:
        local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], 0)
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    # out[r*per + k] is global tile k*world + r  ->  reorder to global order and drop the padding
    out = out.view(world, per, *local.shape[1:]).transpose(0, 1).reshape(world * per, *local.shape[1:])
    return out[:n_tiles].contiguous()
