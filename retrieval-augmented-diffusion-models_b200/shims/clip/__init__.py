"""`clip.tokenize` stand-in (OpenAI CLIP byte-pair tokenizer; the reference vendors the same algorithm in
`rdm/modules/custom_clip/simple_tokenizer.py` / `clip.py:127-143`).  Own implementation of the published BPE procedure; the
merge table (`bpe_simple_vocab_16e6.txt.gz`, a data file of the CLIP release) is NOT shipped here and is looked up at
`$CLIP_BPE_PATH`, then under every `sys.path` root as `rdm/modules/custom_clip/` (drop the file next to this package's mirror of that
directory), `clip/` (an installed CLIP) or the root itself, then under the current directory."""
import gzip
import html
import os
import sys
from functools import lru_cache

import regex as re
import torch

_PAT = re.compile(r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", re.IGNORECASE)


def _find_vocab():
    cands = [os.environ.get("CLIP_BPE_PATH")]
    roots = [os.getcwd()] + list(sys.path)
    for r in roots:
        if r:
            cands += [os.path.join(r, "rdm", "modules", "custom_clip", "bpe_simple_vocab_16e6.txt.gz"), os.path.join(r, "clip", "bpe_simple_vocab_16e6.txt.gz"),
                      os.path.join(r, "bpe_simple_vocab_16e6.txt.gz")]
    for c in cands:
        if c and os.path.isfile(c):
            return c
    raise FileNotFoundError("CLIP BPE merge table bpe_simple_vocab_16e6.txt.gz not found; set CLIP_BPE_PATH")


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs, n = bs[:], 0
    for b in range(256):
        if b not in bs:
            bs.append(b); cs.append(256 + n); n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


class _Tokenizer:
    def __init__(self, path):
        self.byte_encoder = _bytes_to_unicode()
        merges = gzip.open(path).read().decode("utf-8").split("\n")[1:49152 - 256 - 2 + 1]
        merges = [tuple(m.split()) for m in merges]
        vocab = list(self.byte_encoder.values())
        vocab = vocab + [v + "</w>" for v in vocab] + ["".join(m) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = dict(zip(vocab, range(len(vocab))))
        self.ranks = dict(zip(merges, range(len(merges))))
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}

    def _bpe(self, token):
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = set(zip(word[:-1], word[1:]))
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            a, b = best
            out, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == a and word[i + 1] == b:
                    out.append(a + b); i += 2
                else:
                    out.append(word[i]); i += 1
            word = tuple(out)
        self.cache[token] = " ".join(word)
        return self.cache[token]

    def encode(self, text):
        text = html.unescape(html.unescape(text)).strip()
        text = re.sub(r"\s+", " ", text).strip().lower()
        ids = []
        for tok in re.findall(_PAT, text):
            tok = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self._bpe(tok).split(" "))
        return ids


@lru_cache()
def _tokenizer():
    return _Tokenizer(_find_vocab())


def tokenize(texts, context_length=77, truncate=False, cut=False):
    """OpenAI `clip.tokenize` (raises on over-long input unless `truncate`, which re-inserts EOT).  `cut=True` is the behaviour of the
    reference's vendored copy (`rdm/modules/custom_clip/clip.py:127-143`): warn and cut to `context_length` WITHOUT re-inserting EOT."""
    if isinstance(texts, str):
        texts = [texts]
    t = _tokenizer()
    sot, eot = t.encoder["<|startoftext|>"], t.encoder["<|endoftext|>"]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        ids = [sot] + t.encode(text) + [eot]
        if len(ids) > context_length:
            if cut:
                print(f"WARNING: Input of length {len(ids)} is too long for context length {context_length}. Cutting.")
                ids = ids[:context_length]
            elif not truncate:
                raise RuntimeError(f"Input {text} is too long for context length {context_length}")
            else:
                ids = ids[:context_length]; ids[-1] = eot
        result[i, :len(ids)] = torch.tensor(ids)
    return result


def load(name, device="cuda", jit=False, **kw):
    """`clip.load` hands back the reference-shaped (model, preprocess) pair backed by librdm_b200."""
    from rdm.modules.retrievers import load_clip
    return load_clip(name, device=device, jit=jit)
