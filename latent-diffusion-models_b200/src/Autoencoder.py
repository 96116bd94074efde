"""``src.Autoencoder`` shadow."""
from ldm_b200.autoencoder import (  # noqa: F401
    Autoencoder, Encoder, Decoder, ResnetBlock, AttnBlock, UpSample, DownSample, GaussianDistribution,
)
