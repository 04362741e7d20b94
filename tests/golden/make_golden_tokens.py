"""Golden token ids from the REFERENCE's own tokenizer (rdm/modules/custom_clip/simple_tokenizer.py + clip.py:127-143), run in the
build container.  `ftfy` is not installed here; it is stubbed with the identity (it only repairs mojibake, a no-op on these
prompts).  Writes tests/golden/clip_tokens.json.  Run: python tests/golden/make_golden_tokens.py"""
import importlib.util
import json
import os
import sys
import types

REF = "/root/reference/rdm/modules/custom_clip"
PROMPTS = ["a diagram", "a dog", "a cat", "A photo of a  cat, sitting on a sofa!", "an oil painting of the eiffel tower at night; 4k",
           "retrieval-augmented diffusion models (RDM) & friends", "don't stop: it's 100% \"fine\"", ""]


def main():
    sys.modules["ftfy"] = types.SimpleNamespace(fix_text=lambda t: t)
    spec = importlib.util.spec_from_file_location("ref_tok", os.path.join(REF, "simple_tokenizer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tok = mod.SimpleTokenizer()
    sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
    out = {p: [sot] + tok.encode(p) + [eot] for p in PROMPTS}
    json.dump({"sot": sot, "eot": eot, "context_length": 77, "prompts": out}, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "clip_tokens.json"), "w"), indent=0)
    print({k: v[:8] for k, v in out.items()})


if __name__ == "__main__":
    main()
