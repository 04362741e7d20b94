"""Drop-in `rdm` package: the reference's import paths (YAML `target:` strings) backed by librdm_b200.

Only the sampling hot path of CompVis/retrieval-augmented-diffusion-models is provided:
`rdm.models.diffusion.ddpm.MinimalRETRODiffusion`, `rdm.models.diffusion.ddim.DDIMSampler`,
`rdm.modules.diffusionmodules.openaimodel.UNetModel`, `rdm.data.retrieval_dataset.dsetbuilder.DatasetBuilder`.
Training, dataset building and image I/O are out of scope and raise NotImplementedError.
"""
from rdm_b200 import compat as _compat

_compat.install_shims()
