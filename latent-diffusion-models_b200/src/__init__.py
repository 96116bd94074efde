"""Shadow of the reference's ``src`` package for the hot path.

Put ``latent-diffusion-models_b200/`` in front of the reference checkout on ``sys.path`` and the YAML
targets ``src.UNet.UNet`` / ``src.DDPM.Diffusion`` (config_files/*.yaml, resolved by
src/utils.py:48-88) instantiate the B200-native classes; every other ``src.*`` module (Trainer, Config,
utils, data) still comes from the reference checkout through the namespace extension below.
"""
import os as _os
import sys as _sys

# let `src.<anything else>` resolve inside the reference checkout that follows us on sys.path
for _p in _sys.path:
    _cand = _os.path.join(_p, "src")
    if _os.path.isdir(_cand) and _os.path.abspath(_cand) != _os.path.dirname(_os.path.abspath(__file__)):
        if _os.path.exists(_os.path.join(_cand, "DDPM.py")) and _cand not in __path__:
            __path__.append(_cand)
