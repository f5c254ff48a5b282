"""CPU oracle for the MMLF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference algorithm (titus-leistner/mmlf) for the
path SURVEY.md section 8 names: view-index extraction, the disparity Shift
resampler, the FeedForward conv net (forward and a hand-derived backward), the
BASE/UPR/DPP heads, the masked losses (+ gradients), the ESE ensemble reduce
and Adam.  Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package, and only as the checker.
The product package ``mmlf_b200`` never imports it.

Pinning: the reference ships no golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the
build container by ``oracle/gen_golden.py`` (imports ``/root/reference``) and
committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.
"""
from .lf import (view_indices, extract_stacks, shift, shift_taps, ese_shift_values,  # noqa: F401
                 texture_mae, create_mask_texture)
from .net import (FeedForwardOracle, torch_linspace_f32, np_linspace_f32,  # noqa: F401
                  bf16_round, fp16_round, laplacian)
from .losses import (create_mask_margin, reg_to_class, mpi_to_weights, class_to_reg,  # noqa: F401
                     masked_l1, multi_masked_l1, masked_mse, masked_badpix,
                     masked_cross_entropy, improved_uncertainty_l1,
                     improved_multi_uncertainty_l1)
from .ensemble import ensemble_forward, ensemble_reduce  # noqa: F401
from .optim import adam_step  # noqa: F401
from . import augment, metrics  # noqa: F401
