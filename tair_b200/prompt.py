"""Host glue between the text-spotting head and the text encoder: character codec and prompt templates.

Mirrors terediff/dataset/utils.py:18-40 (``CTLABELS`` = the 95 printable ASCII characters 0x20..0x7E, ``decode``
stops at the first index outside the table, ``encode`` pads with 96) and the prompt templates of
terediff/sampler/spaced_sampler.py:309-314.  The recognised indices of a whole batch are brought to the host in a
single copy instead of one ``.cpu()`` per instance (spaced_sampler.py:303-306).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

CTLABELS: List[str] = [chr(c) for c in range(0x20, 0x7F)]
PAD_INDEX = 96
MAX_WORD_LEN = 25


def decode(idxs: Sequence[int]) -> str:
    out = []
    for i in idxs:
        i = int(i)
        if i >= len(CTLABELS):
            break
        out.append(CTLABELS[i])
    return "".join(out)


def encode(word: str) -> List[int]:
    return [CTLABELS.index(word[i]) if i < len(word) else PAD_INDEX for i in range(MAX_WORD_LEN)]


def build_prompt(texts: Sequence[str], style: str = "CAPTION") -> str:
    quoted = ", ".join(f'"{t}"' for t in texts)
    if style == "CAPTION":
        return f"A realistic scene where the texts {quoted} appear clearly on signs, boards, buildings, or other objects."
    if style == "TAG":
        return quoted
    raise ValueError(f"unknown prompt style {style!r}")


def decode_texts(results) -> Tuple[List[List[str]], List[List[np.ndarray]]]:
    """results: one Instances-like object per tile with ``recs`` (k,25) and ``polygons`` (k,32).
    Returns per-tile recognised strings and int32 (16,2) control-point arrays."""
    texts, polys = [], []
    for r in results:
        recs = r.recs.detach().to("cpu", non_blocking=False).numpy() if len(r.recs) else np.zeros((0, MAX_WORD_LEN), np.int64)
        pg = r.polygons.detach().to("cpu").numpy() if len(r.polygons) else np.zeros((0, 32), np.float32)
        texts.append([decode(row) for row in recs])
        polys.append([p.reshape(16, 2).astype(np.int32) for p in pg])
    return texts, polys


_TABLE = np.array([ord(c) for c in CTLABELS] + [0] * (256 - len(CTLABELS)), dtype=np.uint8)


def decode_batch(recs: np.ndarray) -> List[str]:
    """Vectorised ``decode`` of an (n, L) array of character indices: every row is cut at its first index outside the
    table (terediff/dataset/utils.py:21-28) — one numpy pass instead of n*L Python iterations."""
    if recs.size == 0:
        return []
    recs = np.ascontiguousarray(recs)
    valid = recs < len(CTLABELS)
    length = np.where(valid.all(1), recs.shape[1], (~valid).argmax(1))
    codes = _TABLE[np.minimum(recs, 255).astype(np.uint8)]
    raw = codes.tobytes()
    L = recs.shape[1]
    return [raw[i * L:i * L + n].decode("ascii") for i, n in enumerate(length.tolist())]


def texts_and_polys(scores: np.ndarray, polys: np.ndarray, recs: np.ndarray, threshold: float
                    ) -> Tuple[List[List[str]], List[List[np.ndarray]]]:
    """Host half of the detection post-processing: scores (B,Q), polygons (B,Q,32) in pixels, recs (B,Q,L) -> per-tile
    strings and int32 (16,2) polygons of the queries with score >= threshold, in query order
    (transformer_detector.py:133-150, spaced_sampler.py:298-306)."""
    texts, out_polys = [], []
    for b in range(scores.shape[0]):
        keep = scores[b] >= threshold
        texts.append(decode_batch(recs[b][keep]))
        pg = polys[b][keep].astype(np.int32).reshape(-1, polys.shape[2] // 2, 2)
        out_polys.append(list(pg))
    return texts, out_polys
