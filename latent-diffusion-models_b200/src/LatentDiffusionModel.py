"""``src.LatentDiffusionModel`` shadow."""
from ldm_b200.latent import LatentDiffusionModel, DiffusionWrapper  # noqa: F401
