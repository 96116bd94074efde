"""CPU oracle for the DDPM denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline.  The product path
(``latent-diffusion-models_b200/``) never imports this package and fails loudly
when its CUDA library is missing.

Parity pin: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned the other way the task allows: the
unmodified reference is imported from ``/root/reference`` in the build
container by ``oracle/make_golden.py`` and its outputs on seeded inputs are
committed under ``tests/golden/``.  ``tests/test_oracle.py`` checks this
restatement against those fixtures (and against the live reference when it is
present).
"""
from .unet_oracle import unet_forward, init_state_dict, unet_key_shapes  # noqa: F401
from .ddpm_oracle import (  # noqa: F401
    make_schedule,
    q_sample,
    p_sample,
    sample_loop,
)
