"""Oracle (test infrastructure): the Adam update used by the training loop.

The reference calls ``torch.optim.Adam(model.parameters(), lr)`` with default
betas=(0.9, 0.999), eps=1e-8, weight_decay=0 (/root/reference/mmlf/train/cli.py:113-118);
this restates torch's single-tensor Adam update.
"""
import math

import numpy as np


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One update for one tensor; ``step`` is the 1-based step count after the
    increment.  Returns (p, m, v) as float32."""
    g = g.astype(np.float32)
    m = (m + (g - m) * np.float32(1 - beta1)).astype(np.float32)            # lerp
    v = (v * np.float32(beta2) + g * g * np.float32(1 - beta2)).astype(np.float32)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (np.sqrt(v) / np.float32(math.sqrt(bc2)) + np.float32(eps)).astype(np.float32)
    p = (p - np.float32(step_size) * (m / denom)).astype(np.float32)
    return p, m, v
