"""CPU oracle for the Latent2im StyleGAN2 latent-walk hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product path (``latent2im_b200``) never imports this
package and fails loudly when its CUDA library is missing.

Pinning status: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), and its own path is CUDA-only, so it cannot execute in the
CPU build container.  The pins are therefore fixtures produced by running the
UNMODIFIED reference modules on a B200 (``tests/golden/make_golden_ref_gpu.py``
-> ``tests/golden/ref_gpu_*.npz``); ``tests/test_oracle_golden.py`` checks this
oracle against them on CPU.  Until those fixtures exist the oracle is
"parity unpinned".
"""
from .ops import upfirdn2d_ref, fused_bias_act_ref, fused_leaky_relu_ref, make_fir_kernel  # noqa: F401
from .generator import (  # noqa: F401
    GeneratorSpec,
    mapping_ref,
    modulated_conv_ref,
    styled_conv_ref,
    to_rgb_ref,
    generator_forward_ref,
)
from .walks import walk_linear_ref, walk_mlp_ref, walk_nonlinear_ref, z_sample_ref  # noqa: F401
