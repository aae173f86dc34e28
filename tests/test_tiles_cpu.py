"""Host-side tile logic: split parity with the oracle, shard bookkeeping, and the world_size-2 gather order (gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tiles as OT
from tair_b200 import tiles as T


@pytest.mark.parametrize("h,w", [(128, 128), (130, 250), (200, 300), (500, 881)])
def test_split_matches_oracle(h, w):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    mine = [np.asarray(p) for p in T.split_image_with_overlap(img)]
    ref = OT.split_image(img)
    assert len(mine) == len(ref) and all(np.array_equal(a, b) for a, b in zip(mine, ref))
    assert T.tile_grid(h, w) == OT.tile_grid(h, w)


def test_split_grayscale_and_pil():
    from PIL import Image
    img = (np.arange(140 * 150) % 255).astype(np.uint8).reshape(140, 150)
    tiles = T.split_image_with_overlap(Image.fromarray(img))
    assert len(tiles) == 4 and tiles[0].size == (128, 128)
    assert np.array_equal(np.asarray(tiles[3])[:28, :38], img[112:, 112:])


@pytest.mark.parametrize("n,world", [(25, 2), (25, 4), (25, 8), (40, 8), (1, 2), (700, 8)])
def test_shards_partition_the_tile_list(n, world):
    shards = [T.shard_tiles(n, r, world) for r in range(world)]
    assert sorted(i for s in shards for i in s) == list(range(n))
    assert max(len(s) for s in shards) == T.tiles_per_rank(n, world)
    assert all(i % world == r for r, s in enumerate(shards) for i in s)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_tiles, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = T.shard_tiles(n_tiles, rank, world)
        # tile p is filled with the value p so the gathered order is self-describing
        local = torch.stack([torch.full((3, 4, 4), float(p)) for p in mine]) if mine else torch.zeros((0, 3, 4, 4))
        allt = T.gather_tiles(local, n_tiles)
        q.put((rank, allt[:, 0, 0, 0].tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_tiles", [5, 6, 1])
def test_gather_tiles_world2_restores_global_order(n_tiles):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_tiles, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, order in got:
        assert order == [float(i) for i in range(n_tiles)]


def test_bicubic_tables_reproduce_pil_exactly():
    """The host tables behind the GPU tile front-end (tiles.pil_bicubic_coeffs) against PIL itself: the numpy
    restatement of the two fixed-point passes must give PIL's Image.resize(BICUBIC) bytes."""
    from PIL import Image
    from tair_b200.tiles import pil_bicubic_coeffs
    resize_tile_reference = lambda t, n: OT.resize_tile_reference(t, n, pil_bicubic_coeffs)
    rng = np.random.default_rng(3)
    cases = [rng.integers(0, 256, (128, 128, 3), dtype=np.uint8) for _ in range(2)]
    edge = np.zeros((128, 128, 3), np.uint8)
    edge[37:90, 11:77] = 255
    cases.append(edge)
    for t in cases:
        ref = np.asarray(Image.fromarray(t).resize((512, 512), Image.BICUBIC))
        assert np.array_equal(resize_tile_reference(t, 512), ref)
    small = rng.integers(0, 256, (16, 24, 3), dtype=np.uint8)          # non-square, other ratios
    ref = np.asarray(Image.fromarray(small).resize((40, 40), Image.BICUBIC))
    assert np.array_equal(resize_tile_reference(small, 40), ref)
    b, c = pil_bicubic_coeffs(128, 512)
    assert b.shape == (512, 2) and c.shape == (512, 5) and (b[:, 1] <= 5).all()
    assert np.abs(c.sum(1) - (1 << 22)).max() <= 3                       # rows sum to 1.0 in 22-bit fixed point


def test_save_image_matches_to_pil_semantics(tmp_path):
    """pipeline.save_image writes what TF.to_pil_image(...).save() writes (val_patches.py:389-390): mul(255), byte cast."""
    from PIL import Image
    from tair_b200.pipeline import save_image
    x = torch.rand((1, 3, 9, 7), generator=torch.Generator().manual_seed(0))
    p = str(tmp_path / "restored.png")
    save_image(x, p)
    assert np.array_equal(np.asarray(Image.open(p)), x[0].mul(255).byte().permute(1, 2, 0).numpy())
