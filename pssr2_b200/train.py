"""Crappifier approximation (pssr/train.py:324-386): the objective that compares the noise profile of an artificial crappifier with
ground-truth low-resolution images, evaluated on the device.

Per sample the reference downsamples the HR tile with Pillow, runs the candidate crappifier on it, and compares two things with the real
LR tile: the histogram of (image - downsampled HR) over np.arange(-256, 256) and the mean of that profile.  Here the downscale is the
Pillow-exact resize kernel, the crappifier is its device noise chain and the two histograms / sums come from ``pssr_profile_hist``;
nothing but a 511-bin histogram per image returns to the host.  ``approximate_crappifier`` hands the objective to scikit-optimize's
``gp_minimize`` exactly like the reference (the package is an optional dependency).  Model training itself is out of scope.
"""
import random

import numpy as np
import torch

from . import ops
from .crappifiers import Crappifier, _DeviceCrappifier, _fresh_seed, run_noise_chain


class _Crappifier_Objective():
    def __init__(self, crappifier, dataset, n_samples: int):
        self.crappifier = crappifier
        self.dataset = dataset
        self.n_samples = n_samples

    def sample(self, params):
        sample_idx = list(range(len(self.dataset)))
        random.shuffle(sample_idx)
        crap = self.crappifier(*params)
        metrics = []
        for idx in sample_idx[:self.n_samples]:
            hr, lr = self.dataset[idx]                                   # float32 device tensors [f, H, W] / [f, h, w]
            hr8, lr8 = hr.to(torch.uint8), lr.to(torch.uint8)             # np.asarray(..., dtype=np.uint8): truncation
            scale = hr8.shape[-1] // lr8.shape[-1]
            ds_hr = ops.resize_bilinear(hr8, scale)                       # Pillow BILINEAR to the LR size (train.py:365)
            specs = crap.noise_specs() if isinstance(crap, _DeviceCrappifier) else None
            if specs is not None and len(specs) <= 4:
                lr_hat = run_noise_chain(ds_hr.to(torch.float32), specs, crap.clip_between, _fresh_seed())     # float64, on the device
            else:                                                        # a custom host crappifier (any callable)
                out = crap.crappify(ds_hr.cpu().numpy()) if isinstance(crap, Crappifier) else crap(ds_hr.cpu().numpy())
                lr_hat = torch.as_tensor(np.asarray(out, dtype=np.float64)).to(ds_hr.device)
            pred_dist, pred_sum = ops.profile_hist(lr_hat, ds_hr)
            target_dist, target_sum = ops.profile_hist(lr8, ds_hr)
            both = torch.cat([pred_dist.double(), target_dist.double(), pred_sum, target_sum]).cpu().numpy()
            pd, td, ps, ts = both[:511], both[511:1022], both[1022], both[1023]
            n = float(lr8.numel())
            # np.mean over the 511 bins of the squared count difference, scaled by the image area (train.py:376-380)
            dist_error = float(np.mean((td - pd) ** 2)) / (lr8.shape[-1] ** 2)
            value_error = abs(ts / n - ps / n)
            metrics.append(dist_error + value_error)
        return sum(metrics) / len(metrics)


def approximate_crappifier(crappifier, space, dataset, max_images=None, opt_kwargs=None):
    r"""Approximates :class:`Crappifier` parameters from ground truth paired images with Bayesian optimisation
    (pssr/train.py:324-346).  Same arguments as the reference; needs scikit-optimize."""
    try:
        from skopt import gp_minimize
    except ImportError as e:
        raise ImportError("approximate_crappifier needs scikit-optimize (skopt.gp_minimize), as the reference does") from e
    space = [space] if type(space) is not list else space
    n_samples = len(dataset) if max_images is None else min(max_images, len(dataset))
    opt_kwargs = {} if opt_kwargs is None else opt_kwargs
    objective = _Crappifier_Objective(crappifier, dataset, n_samples).sample
    return gp_minimize(objective, space, **opt_kwargs)
