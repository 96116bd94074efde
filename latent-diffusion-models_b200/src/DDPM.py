"""``src.DDPM`` shadow: the YAML target ``src.DDPM.Diffusion`` resolves to the B200-native class."""
from ldm_b200.ddpm import Diffusion  # noqa: F401
