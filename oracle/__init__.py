"""CPU oracle for the retrieval-augmented diffusion sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  The product path (``retrieval-augmented-diffusion-models_b200/``)
never imports this package and fails loudly when its CUDA library is missing.

PINNING.  The reference ships no tests, fixtures or golden vectors
and its modules do not import as they are (``ldm``, ``scann``, ``clip``,
``omegaconf``, ``pytorch_lightning``, ``kornia`` are absent and un-vendored,
SURVEY.md §8c).  The restatements here follow

* ``rdm/modules/diffusionmodules/openaimodel.py:17-33,66-317,335-371`` (U-Net),
* ``rdm/modules/attention.py:16-17,20-74,77-96,122-196`` (SpatialTransformer), ``:199-272`` (RARM decoder),
* ``rdm/models/diffusion/ddim.py:27-56,143-215,218-268`` (DDIM sampler),
* ``rdm/models/autoregression/transformer.py:224-270`` (RARM sampling loop),
* ``rdm/data/retrieval_dataset/dsetbuilder.py:478-518,574`` + ``rdm/models/diffusion/ddpm.py:897-921`` (kNN),
* the public ``latent-diffusion@main`` pieces those files import,

and are PINNED to outputs of the reference's own code: ``tests/golden/make_golden_ref.py``
runs the reference's UNetModel / SpatialTransformer / DDIMSampler /
MinimalRETRODiffusion.sample_from_rdata / sample_with_query / RetrievalPatchTransformer
unmodified from /root/reference on CPU (over small stand-ins for the un-vendored
``ldm`` functions, ``tests/golden/ref_stubs.py``) and commits the outputs
(``tests/golden/ref_*.npz``); ``tests/golden/make_golden.py`` does the same for the
vendored CLIP.  ``tests/test_oracle_ref_golden.py``, ``test_oracle_rarm.py`` and
``test_oracle_clip.py`` check the oracle against them; ``tests/test_oracle.py`` adds the
known answers printed in the repository (400,920,579 U-Net parameters in 688 tensors,
``scripts/demo_rdm.ipynb:128``; head counts 12/18/30; the DDIM-100 grid 1,11,…,991).

PARITY UNPINNED remains for two pieces: the kNN result (ScaNN 1.2.4 is an approximate CPU
searcher that is not installed; ``oracle/knn_ref.c`` defines the exact semantics it
approximates) and the first-stage decoder (``oracle/vqdecoder.py`` restates ldm's
un-vendored ``Decoder``).
"""
