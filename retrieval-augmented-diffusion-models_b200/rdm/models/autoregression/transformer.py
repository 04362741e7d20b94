"""`rdm.models.autoregression.transformer.LatentImageRETRO` -- the sampling side of the reference's RARM model
(`rdm/models/autoregression/transformer.py:122-391`; base classes `LatentCrossTransformer` :24-119 and taming's
`Net2NetTransformer`, which is not vendored: the pieces used when sampling -- the SOS conditioning of `__is_unconditional__` models,
`top_k_logits`, `decode_to_img` -- are restated from the published taming-transformers source; the VQGAN-f16 first stage is
`shims/taming/models/vqgan.py` over the device decoder when taming itself is not installed).

`sample` (:224-270) runs on the device decoder of librdm_b200 (csrc/rarm.cu): key/value caches instead of re-evaluating the growing
prefix, guidance on the logits, top-k filter and the draw fused into one kernel, one CUDA-graph replay per position.  The draw uses
uniforms from `torch.rand` (the global generator `seed_everything` seeds), so runs are reproducible; `torch.multinomial`'s own
stream cannot be reproduced by a device kernel (same distribution, see oracle/rarm.py).  Training (`forward`, `shared_step`,
`configure_optimizers`, `log_images`) is out of scope."""
import os
import pickle
import time

import numpy as np
import torch
import torch.nn as nn

from ldm.util import instantiate_from_config

from rdm.modules.encoders.nn_encoders import IdentityEncoder
from rdm.util import SampleLogs
from rdm_b200.knn import search_raw


def disabled_train(self, mode=True):
    return self


class SOSProvider(nn.Module):
    """taming `SOSProvider`: the conditioning of unconditional models is one start-of-sequence token per example."""

    def __init__(self, sos_token, quantize_interface=True):
        super().__init__()
        self.sos_token, self.quantize_interface = sos_token, quantize_interface

    def encode(self, x):
        c = (torch.ones(x.shape[0], 1) * self.sos_token).long().to(x.device)
        return (c, None, [None, None, c]) if self.quantize_interface else c


class LatentImageRETRO(nn.Module):
    def __init__(self, nn_encoder_cfg, nn_key, mask_token, p_mask_max=0., nn_reshaper_cfg=None, nn_memory=None, retrieval_cfg=None,
                 scheduler_config=None,
                 # taming Net2NetTransformer keys
                 transformer_config=None, first_stage_config=None, cond_stage_config="__is_unconditional__", permuter_config=None, ckpt_path=None,
                 ignore_keys=(), first_stage_key="image", cond_stage_key="depth", downsample_cond_size=-1, pkeep=1.0, sos_token=0, unconditional=False,
                 **unused):
        super().__init__()
        assert pkeep == 1.0, 'currently only supporting pkeep=1.0'
        if cond_stage_config != "__is_unconditional__" or permuter_config is not None:
            raise NotImplementedError("the shipped RARM configs are unconditional (SOS token) with the identity permuter")
        self.be_unconditional, self.first_stage_key, self.cond_stage_key = True, first_stage_key, first_stage_key
        print(f"Using no cond stage. Assuming the training is intended to be unconditional. Prepending {sos_token} as a sos token.")
        self.cond_stage_model = SOSProvider(sos_token)
        try:                                                                   # taming VQGAN-f16 (decode only): used when importable
            self.first_stage_model = instantiate_from_config(first_stage_config).eval() if first_stage_config else None
        except ImportError as e:
            print(f"first stage {first_stage_config.get('target')} is not importable ({e}); sampling returns token ids, decode_to_img raises")
            self.first_stage_model = None
        self.transformer = instantiate_from_config(transformer_config)
        self.register_buffer("sos_token", torch.LongTensor([sos_token]))
        self.register_buffer("mask_token", torch.LongTensor([mask_token]))
        self.p_mask_max, self.nn_key, self.pkeep = p_mask_max, nn_key, pkeep
        self.nn_encoder = instantiate_from_config(nn_encoder_cfg).eval()
        self.nn_encoder.train = disabled_train
        self.nn_reshaper = instantiate_from_config(nn_reshaper_cfg) if nn_reshaper_cfg is not None else torch.nn.Identity()
        self.use_memory = nn_memory is not None
        if self.use_memory:
            assert os.path.isfile(nn_memory) and nn_memory.endswith('.p')
            with open(nn_memory, 'rb') as f:
                nn_data = pickle.load(f)
            self.register_buffer('nn_memory', torch.tensor(nn_data['nn_memory'], dtype=torch.int), persistent=False)
            print(f'Loaded nn_memory of size {self.nn_memory.shape[0]}')
            self.id_count = nn_data['id_count']
        self.retriever = None
        self.init_retriever(retrieval_cfg)
        if ckpt_path is not None:
            sd = torch.load(ckpt_path, map_location="cpu")["state_dict"]
            for k in [k for k in sd if any(k.startswith(ik) for ik in ignore_keys)]:
                del sd[k]
            self.load_state_dict(sd, strict=False)

    @property
    def device(self):
        return self.sos_token.device

    def load_state_dict(self, state_dict, strict=True):
        if self.first_stage_model is None:                                     # no container for the first-stage tensors: ignore them
            state_dict = {k: v for k, v in state_dict.items() if not k.startswith("first_stage_model.")}
        return super().load_state_dict(state_dict, strict=strict)

    def init_retriever(self, cfg):
        if not cfg:
            self.retriever = None
            return
        self.retriever = instantiate_from_config(cfg)
        self.retriever.train = disabled_train

    def train_searcher(self):
        print("training searcher...")
        self.retriever.train_searcher(device=self.device if self.device.type == "cuda" else None)
        print("done training searcher")

    @torch.no_grad()
    def encode_nns(self, nns):
        return self.nn_encoder.encode(nns)

    @torch.no_grad()
    def encode_to_c(self, c):
        quant_c, _, [_, _, indices] = self.cond_stage_model.encode(c)
        if len(indices.shape) > 2:
            indices = indices.view(c.shape[0], -1)
        return quant_c, indices

    def top_k_logits(self, logits, k):
        v, ix = torch.topk(logits, k)
        out = logits.clone()
        out[out < v[..., [-1]]] = -float('Inf')
        return out

    # ---- sampling (transformer.py:224-270) ---------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, x, r, c, steps, temperature=1.0, sample=False, top_k=None, guidance_scale=1.0, callback=lambda k: None, **kwargs):
        assert not self.transformer.training
        x = torch.cat((c, x), 1)
        if guidance_scale > 1.0:
            r = torch.cat((r, torch.zeros_like(r)), dim=0)
        eng = self.transformer.engine(r.device)
        eng.set_context(r)
        uniforms = torch.rand((steps, x.shape[0]), device=r.device) if sample else None
        out = eng.sample(x, steps, temperature=temperature, top_k=top_k, guidance_scale=guidance_scale, uniforms=uniforms)
        for k in range(steps):                                                 # the loop runs on the device; progress callbacks fire afterwards
            callback(k)
        return out[:, c.shape[1]:]

    @torch.no_grad()
    def decode_to_img(self, index, zshape):
        if self.first_stage_model is None:
            raise NotImplementedError("this model was built without a first stage (first_stage_config missing or not importable)")
        bhwc = (zshape[0], zshape[2], zshape[3], zshape[1])
        quant_z = self.first_stage_model.quantize.get_codebook_entry(index.reshape(-1), shape=bhwc)
        return self.first_stage_model.decode(quant_z)

    @torch.no_grad()
    def sampling_util(self, steps, z_start, r, c, temperature, top_k, zshape, callback=None, top_p=1., **kwargs):
        assert top_p == 1., 'not yet implemented'
        t1 = time.time()
        index_sample = self.sample(z_start, r, c, steps=steps, temperature=temperature if temperature is not None else 1.0, sample=True,
                                   top_k=top_k if top_k is not None else 100, callback=callback if callback is not None else lambda k: None, **kwargs)
        if r.is_cuda:
            torch.cuda.synchronize(r.device)
        if not hasattr(self, "sampling_time"):
            self.sampling_time = time.time() - t1
            print(f"Full sampling takes about {self.sampling_time:.2f} seconds.")
        self.last_index_sample = index_sample
        return self.decode_to_img(index_sample, zshape) if self.first_stage_model is not None else index_sample

    def _searcher(self):
        if self.retriever.searcher is None:
            self.train_searcher()
        return self.retriever.searcher

    @torch.no_grad()
    def sample_from_rdata(self, N, cond=None, return_nns=False, use_weights=False, qids=None, k_nn=None, memsize=100, verbose=False, top_k=256,
                          temperature=1.0, code_side_len=16, z_dimensionality=256, pre_loaded_patches=None, nn_embeddings=None, query_embeddings=None,
                          **kwargs):
        """transformer.py:296-391 with the retrieval on the device: exact kNN + gather (librdm_b200) instead of ScaNN + host gather."""
        if cond is not None:
            raise NotImplementedError()
        if return_nns or pre_loaded_patches is not None or not isinstance(self.nn_encoder, IdentityEncoder):
            raise NotImplementedError("neighbour image patches / VQ neighbour encoders need the patch dataset, which is outside the sampling hot path")
        if k_nn is None:
            k_nn = self.k_nn
        out = SampleLogs()
        if nn_embeddings is None:
            searcher = self._searcher()
            if query_embeddings is None:
                qids = self.get_qids(memsize, N, qids=qids, use_weights=use_weights, verbose=verbose)
                out["qids"] = qids
                q = searcher.gather_device(torch.as_tensor(np.asarray(qids), dtype=torch.int64, device=self.device))      # :320
            else:
                q = torch.as_tensor(np.asarray(query_embeddings), dtype=torch.float32).to(self.device)
            nns, _ = search_raw(searcher, q, k_nn)                                                                         # :327-329
            retro_cond = searcher.gather_device(nns)                                                                       # :342
            out.extras["nns"] = nns
        else:
            retro_cond = nn_embeddings.to(self.device, torch.float32)
        _, c = self.encode_to_c(torch.zeros((N, 0)))                            # SOSProvider conditioning (:377-379)
        c = c.to(retro_cond.device)
        z_shape = (N, z_dimensionality, code_side_len, code_side_len)
        steps = code_side_len ** 2
        z_start = torch.zeros((N, 0), device=retro_cond.device, dtype=torch.long)
        out["samples_with_sampled_nns"] = self.sampling_util(steps, z_start, retro_cond, c, temperature, top_k, z_shape, **kwargs)
        out.extras["sampled_indices"] = self.last_index_sample
        return out

    def get_qids(self, memsize, N, qids=None, use_weights=False, verbose=False):
        """transformer.py:394-420 (NumPy global RNG, host)."""
        if isinstance(memsize, float):
            assert 0 < memsize <= 1., 'Require memsize in (0,1]'
            memsize = int(memsize * self.nn_memory.shape[0])
        if qids is None:
            if self.use_memory:
                memsize = min(memsize, self.nn_memory.shape[0])
                print(f'Top-M Sampling with memory size {memsize}')
                nn_mem = self.nn_memory.detach().cpu().numpy()[:memsize]
                ps = None
                if use_weights:
                    freqs = np.asarray([self.id_count[int(id_)] for id_ in nn_mem])
                    ps = freqs / freqs.sum(keepdims=True)
                qids = np.random.choice(nn_mem, size=N, p=ps)
            else:
                print('Randomly sampling retrieval database entries')
                qids = np.random.choice(getattr(self.retriever, 'num_rows', None) or len(self.retriever.data_pool['embedding']), size=N)   # global row count, also when row-sharded
        else:
            assert qids.shape[0] == N
        if verbose:
            print(f'Sampled entries are {qids}')
        return qids

    def _training_only(self, *a, **k):
        raise NotImplementedError("training is outside the sampling hot path (SURVEY.md section 2)")

    forward = shared_step = training_step = validation_step = configure_optimizers = log_images = _training_only
