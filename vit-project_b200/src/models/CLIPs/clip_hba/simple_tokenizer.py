"""Byte-pair-encoding tokenizer of the published CLIP checkpoints (the `clip.tokenize` the reference calls at
NEW:282 resolves to it in the un-vendored CLIP-HBA-Official package).

Restated from the published algorithm: bytes are mapped to printable unicode characters, a word is a tuple
of such characters with `</w>` appended to the last one, adjacent pairs are merged in the order of the 48,894
merge rules of `bpe_simple_vocab_16e6.txt.gz`, ids follow the vocabulary order 256 byte symbols, the same 256
with `</w>`, the merges, `<|startoftext|>` (49406), `<|endoftext|>` (49407).

The merge table is data (1.3 MB), not code, and there is no network here: it is looked up at
  $HBA_BPE_VOCAB, <this directory>/bpe_simple_vocab_16e6.txt.gz, ~/.cache/clip/bpe_simple_vocab_16e6.txt.gz
`load()` returns None when none exists; `clip.tokenize` then refuses to tokenise for a real checkpoint.
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache

import regex as re

VOCAB_FILE = "bpe_simple_vocab_16e6.txt.gz"
_N_MERGES = 49152 - 256 - 2


def vocab_candidates():
    here = os.path.dirname(os.path.abspath(__file__))
    env = os.environ.get("HBA_BPE_VOCAB")
    return ([env] if env else []) + [os.path.join(here, VOCAB_FILE),
                                     os.path.join(os.path.expanduser("~/.cache/clip"), VOCAB_FILE)]


@lru_cache()
def byte_symbols():
    """byte value -> printable unicode character (printable latin-1 bytes map to themselves, the other 68
    to code points from 256 upwards)."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    # vocabulary order of the published table: the printable bytes first, then the remapped ones
    order = keep + [b for b in range(256) if b not in keep]
    return table, [table[b] for b in order]


class SimpleTokenizer:
    def __init__(self, vocab_path):
        with gzip.open(vocab_path, "rt", encoding="utf-8") as f:
            lines = f.read().split("\n")
        merges = [tuple(line.split()) for line in lines[1:_N_MERGES + 1]]
        self.byte_map, symbols = byte_symbols()
        vocab = list(symbols) + [s + "</w>" for s in symbols] + ["".join(m) for m in merges]
        vocab += ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.ranks = {m: i for i, m in enumerate(merges)}
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}
        self.pattern = re.compile(
            r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
            re.IGNORECASE)
        self.sot, self.eot = self.encoder["<|startoftext|>"], self.encoder["<|endoftext|>"]

    def _merge_word(self, token):
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = {(word[i], word[i + 1]) for i in range(len(word) - 1)}
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            a, b = best
            merged, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == a and word[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = tuple(merged)
        out = " ".join(word)
        self.cache[token] = out
        return out

    @staticmethod
    def clean(text):
        try:                       # the published tokenizer repairs mojibake first; identity on plain text
            import ftfy
            text = ftfy.fix_text(text)
        except ImportError:
            pass
        text = html.unescape(html.unescape(text)).strip()
        return re.sub(r"\s+", " ", text).strip().lower()

    def encode(self, text):
        ids = []
        for token in self.pattern.findall(self.clean(text)):
            token = "".join(self.byte_map[b] for b in token.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self._merge_word(token).split(" "))
        return ids


@lru_cache()
def load():
    for path in vocab_candidates():
        if path and os.path.exists(path):
            return SimpleTokenizer(path)
    return None
