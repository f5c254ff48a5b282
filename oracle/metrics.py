"""Oracle (test infrastructure): the distribution metrics of the validation script.

Restates /root/reference/mmlf/validate/cli.py:74-115 (Laplace / Laplace-mixture posteriors integrated over the 108
disparity bins through the Laplace CDF), :130-137 (mean_to_discrete), :165-171 (multimodal_mask), :174-187
(kl_divergence) and :51-70 (nll_discrete) as PURE functions: the reference versions add epsilon to and re-normalise
their arguments in place, and validate.main calls kl_divergence three times on the same arrays (:323-325), so the second
and third results depend on that mutation -- here every function returns the mutated arrays next to its value so a
caller can replay the sequence.
"""
import numpy as np

EPS = 0.00001


def cdf_laplace(disp, mean, var):
    """validate/cli.py:74-88."""
    z = (disp - mean) / var
    return np.where(disp < mean, np.exp(z) / 2, 1 - np.exp(-z) / 2)


def laplace_to_discrete(n_bins, x_min, x_max, mean, logvar):
    """validate/cli.py:91-104.  mean / logvar (B, H, W) float32 -> (B, n_bins, H, W) float64: CDF differences over
    n_bins + 1 edges from x_min - step/2 to x_max + step/2.  var = exp(logvar) stays float32 like the input."""
    step = (x_max - x_min) / n_bins
    edges = np.linspace(x_min - step / 2.0, x_max + step / 2.0, n_bins + 1).reshape(1, -1, 1, 1)
    cdf = cdf_laplace(edges, mean[:, None], np.exp(logvar[:, None]))
    return cdf[:, 1:] - cdf[:, :-1]


def lmm_to_discrete(n_bins, x_min, x_max, means, logvars):
    """validate/cli.py:107-118: average of the members' discretised Laplacians; means / logvars (K, B, H, W)."""
    out = np.zeros((means.shape[1], n_bins, means.shape[2], means.shape[3]))
    for i in range(means.shape[0]):
        out += laplace_to_discrete(n_bins, x_min, x_max, means[i], logvars[i])
    return out / means.shape[0]


def mean_to_discrete(n_bins, x_min, x_max, mean):
    """validate/cli.py:121-137."""
    step = (x_max - x_min) / n_bins
    bins = np.linspace(x_min, x_max, n_bins).reshape(1, -1, 1, 1)
    return (np.abs(bins - mean[:, None]) < step / 2.0).astype(float)


def multimodal_mask(mpi, threshold=0.3):
    """validate/cli.py:165-171."""
    return (np.sum(mpi[:, :, 3] > threshold, 1) > 1).astype(float)


def kl_divergence(dist, dist_gt, mask=None):
    """validate/cli.py:174-187 for batch size 1 (the script's ``dist /= np.sum(dist, 1)`` only broadcasts for B = 1).
    Returns (value, dist', dist_gt') with the primed arrays = the reference's in-place results."""
    dist = dist + EPS
    dist_gt = dist_gt + EPS
    dist = dist / np.sum(dist, 1, keepdims=True)
    dist_gt = dist_gt / np.sum(dist_gt, 1, keepdims=True)
    kld = np.sum(dist_gt * np.log(dist_gt / dist), 1)
    val = np.mean(kld) if mask is None else np.sum(kld * mask) / np.sum(mask)
    return val, dist, dist_gt


def nll_discrete(weights, posterior, mask=None):
    """validate/cli.py:51-70 (the 7.0 is the disparity range hard-coded there).  Returns (value, weights', posterior')."""
    weights = weights + EPS
    posterior = posterior + EPS
    weights = weights / np.sum(weights, 1, keepdims=True)
    posterior = posterior / (np.sum(posterior, 1, keepdims=True) * 7.0)
    nllh = np.sum(weights * -np.log(posterior), axis=1)
    val = np.mean(nllh) if mask is None else np.sum(nllh * mask) / np.sum(mask)
    return val, weights, posterior
