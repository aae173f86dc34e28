"""Host-side CLIP byte-pair tokenizer (row 21 of SURVEY.md §8a: prompt string -> (B,77) token ids for the text encoder).

Behaviour follows the tokenizer the reference calls through ``open_clip.tokenize``
(terediff/model/open_clip/tokenizer.py:72-188): byte->printable-unicode alphabet, 48 894 ranked merges read from
``bpe_simple_vocab_16e6.txt.gz`` (line 0 is a header), vocabulary = 256 byte symbols + 256 word-final symbols + one
symbol per merge + ``<start_of_text>``/``<end_of_text>``; text is HTML-unescaped twice, whitespace-collapsed and
lower-cased; ids are wrapped in start/end markers, truncated to the context length (last id forced to end-of-text)
and zero padded.  (The reference also runs ``ftfy.fix_text`` first; ftfy is not in this image, and it is the identity
on the ASCII prompts the sampler builds from the 96-character recognition alphabet.)

The merge table is a 1.3 MB data file that belongs to OpenAI CLIP and is not redistributed here: pass its path, or
set ``TAIR_BPE_VOCAB``.  Without it construction fails loudly.
"""
from __future__ import annotations

import gzip
import html
import os
from typing import Dict, Iterable, List, Sequence, Tuple, Union

import regex
import torch

N_MERGES = 49152 - 256 - 2
SOT, EOT = "<start_of_text>", "<end_of_text>"
WORD_END = "</w>"


def _byte_alphabet() -> Dict[int, str]:
    """Printable stand-in for each of the 256 byte values: bytes that already are visible latin-1 characters keep their
    code point, the remaining 68 are moved to U+0100.. in ascending byte order."""
    visible = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    table, spill = {}, 0
    for b in sorted(visible):
        table[b] = chr(b)
    for b in range(256):
        if b not in visible:
            table[b] = chr(256 + spill)
            spill += 1
    return table


def _alphabet_in_vocab_order() -> List[str]:
    # vocabulary order is "visible bytes first (ascending), then the relocated ones": ids 0..255
    t = _byte_alphabet()
    visible = [b for b in range(256) if ord(t[b]) < 256]
    hidden = [b for b in range(256) if ord(t[b]) >= 256]
    return [t[b] for b in visible + hidden]


class BPETokenizer:
    def __init__(self, bpe_path: Union[str, None] = None):
        bpe_path = bpe_path or os.environ.get("TAIR_BPE_VOCAB")
        if not bpe_path or not os.path.exists(bpe_path):
            raise FileNotFoundError("BPETokenizer needs the CLIP merge table bpe_simple_vocab_16e6.txt.gz: pass its "
                                    "path or set TAIR_BPE_VOCAB (the file is not shipped with tair_b200)")
        opener = gzip.open if bpe_path.endswith(".gz") else open
        with opener(bpe_path, "rb") as f:
            lines = f.read().decode("utf-8").split("\n")
        pairs = [tuple(ln.split()) for ln in lines[1:N_MERGES + 1]]
        pairs = [p for p in pairs if len(p) == 2]
        self.byte_sym = _byte_alphabet()
        self.sym_byte = {s: b for b, s in self.byte_sym.items()}
        alphabet = _alphabet_in_vocab_order()
        symbols = alphabet + [s + WORD_END for s in alphabet] + [a + b for a, b in pairs] + [SOT, EOT]
        self.encoder: Dict[str, int] = {s: i for i, s in enumerate(symbols)}
        self.decoder: Dict[int, str] = {i: s for s, i in self.encoder.items()}
        self.rank: Dict[Tuple[str, str], int] = {p: i for i, p in enumerate(pairs)}
        self.sot_id, self.eot_id = self.encoder[SOT], self.encoder[EOT]
        self.vocab_size = len(self.encoder)
        self._split = regex.compile(SOT + "|" + EOT + r"|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                                    regex.IGNORECASE)
        self._memo: Dict[str, List[int]] = {SOT: [self.sot_id], EOT: [self.eot_id]}

    # ------------------------------------------------------------------ merging
    def _merge_word(self, word: str) -> List[int]:
        """Greedy lowest-rank-first merging of one pre-token (already mapped to the byte alphabet)."""
        hit = self._memo.get(word)
        if hit is not None:
            return hit
        parts = list(word[:-1]) + [word[-1] + WORD_END]
        while len(parts) > 1:
            best, best_rank = None, None
            for a, b in zip(parts, parts[1:]):
                r = self.rank.get((a, b))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = (a, b), r
            if best is None:
                break
            merged, k = [], 0
            while k < len(parts):  # left-to-right, non-overlapping replacement of every occurrence
                if k + 1 < len(parts) and parts[k] == best[0] and parts[k + 1] == best[1]:
                    merged.append(parts[k] + parts[k + 1])
                    k += 2
                else:
                    merged.append(parts[k])
                    k += 1
            parts = merged
        ids = [self.encoder[p] for p in parts]
        self._memo[word] = ids
        return ids

    @staticmethod
    def _clean(text: str) -> str:
        text = html.unescape(html.unescape(text)).strip()
        return regex.sub(r"\s+", " ", text).strip().lower()

    def encode(self, text: str, max_tokens: int = 0) -> List[int]:
        """``max_tokens`` > 0 stops once that many tokens exist (pre-tokens are independent, so the prefix is exact): the
        sampler's prompts list up to 100 recognised strings per tile per step and only the first 75 tokens survive."""
        out: List[int] = []
        for m in self._split.finditer(self._clean(text)):
            out.extend(self._merge_word("".join(self.byte_sym[b] for b in m.group(0).encode("utf-8"))))
            if max_tokens and len(out) >= max_tokens:
                break
        return out

    def decode(self, ids: Iterable[int]) -> str:
        chars = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(self.sym_byte[c] for c in chars).decode("utf-8", errors="replace").replace(WORD_END, " ")

    def __call__(self, texts: Union[str, Sequence[str]], context_length: int = 77) -> torch.Tensor:
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros((len(texts), context_length), dtype=torch.long)
        for i, t in enumerate(texts):
            ids = [self.sot_id] + self.encode(t, context_length) + [self.eot_id]
            if len(ids) > context_length:
                ids = ids[:context_length]
                ids[-1] = self.eot_id
            out[i, :len(ids)] = torch.tensor(ids)
        return out
