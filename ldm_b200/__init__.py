"""Import alias: ``latent-diffusion-models_b200`` is not a valid Python identifier, so this package
re-exports it under the importable name ``ldm_b200`` (same module objects, one copy of the code)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "latent-diffusion-models_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
