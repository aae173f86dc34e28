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
