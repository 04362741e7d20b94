"""`rdm.modules.custom_clip.clip` -- the import path of the reference's vendored CLIP front end (`rdm/modules/custom_clip/clip.py`).
Only what the sampling path touches lives here: `tokenize` with the vendored copy's truncation rule (`:127-143`: over-long captions are
cut to the context length with a warning, EOT is NOT re-inserted) and `load`, which hands back the device executor (csrc/clip.cu)."""
from rdm_b200 import compat as _compat

_compat.install_shims()           # makes the `clip` stand-in importable when OpenAI CLIP is not installed


def tokenize(texts, context_length=77):
    from clip import tokenize as _tok
    try:
        return _tok(texts, context_length=context_length, cut=True)
    except TypeError:                     # a real OpenAI `clip` package is installed: reproduce the cut on its ids
        import torch
        from clip.simple_tokenizer import SimpleTokenizer
        t = SimpleTokenizer()
        texts = [texts] if isinstance(texts, str) else texts
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, text in enumerate(texts):
            ids = [t.encoder["<|startoftext|>"]] + t.encode(text) + [t.encoder["<|endoftext|>"]]
            if len(ids) > context_length:
                print(f"WARNING: Input of length {len(ids)} is too long for context length {context_length}. Cutting.")
                ids = ids[:context_length]
            out[i, :len(ids)] = torch.tensor(ids)
        return out


def load(name, device="cuda", jit=False, **kw):
    from rdm.modules.retrievers import load_clip
    return load_clip(name, device=device, jit=jit)
