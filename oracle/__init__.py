"""CPU oracle for the retrieval-augmented diffusion sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  The product path (``retrieval-augmented-diffusion-models_b200/``)
never imports this package and fails loudly when its CUDA library is missing.

PARITY UNPINNED by the reference: the reference ships no tests, fixtures or
golden vectors, its own modules cannot be imported here (``ldm``,
``scann``, ``clip``, ``omegaconf``, ``pytorch_lightning``, ``kornia`` are absent
and un-vendored), and its kNN backend (ScaNN 1.2.4) is an
*approximate* CPU searcher.  The oracle is therefore a restatement of

* ``rdm/modules/diffusionmodules/openaimodel.py:17-33,66-317,335-371`` (U-Net),
* ``rdm/modules/attention.py:16-17,20-74,77-96,122-196`` (SpatialTransformer),
* ``rdm/models/diffusion/ddim.py:27-56,143-215,218-268`` (DDIM sampler),
* ``rdm/data/retrieval_dataset/dsetbuilder.py:478-518,574`` + ``rdm/models/diffusion/ddpm.py:897-921`` (kNN),
* the public ``latent-diffusion@main`` pieces those files import,

pinned only by the known answers the reference repo does contain: the printed
U-Net size (400.92 M parameters, ``scripts/demo_rdm.ipynb:128``; exact count
400,920,579 in 688 tensors), the per-level head counts 12/18/30 at d_head 32
(``scripts/demo_rdm.ipynb:112-127``), the DDIM-100 timestep grid 1,11,…,991 and
the conditioning shape ``[8, 4, 512]``.  ``tests/test_oracle.py`` asserts them.
The vendored CLIP (``rdm/modules/custom_clip/model.py``) does import on CPU and
is used by ``tests/golden/make_golden.py`` to pin the CLIP restatement.
"""
