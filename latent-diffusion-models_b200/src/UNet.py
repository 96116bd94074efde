"""``src.UNet`` shadow: the YAML target ``src.UNet.UNet`` resolves to the B200-native class."""
from ldm_b200.unet import UNet  # noqa: F401
