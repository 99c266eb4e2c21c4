"""
Acquisition functions over the dense grid: confidence bound, expected improvement, probability of
improvement.  Same signatures and return values as the reference's gpim/gpbayes/acqfunc.py
(:11-92) -- ``(acq, (mean, sd))`` as numpy arrays -- but the element-wise sweep runs on the device
(``gpg_acq_sweep``) on the prediction that is still resident there.
"""
import numpy as np
import torch

from .._lib import ACQ_IDS


def _sweep(gpmodel, acq_name, mean, sd, mu_best=0.0, xi=0.01, alpha=0.0, beta=1.0):
    eng = gpmodel.model.engine
    dev = getattr(gpmodel, "_last_pred_device", None)
    if dev is not None and dev[0].numel() == mean.size:
        mean_d, sd_d = dev
    else:
        mean_d = torch.as_tensor(mean.ravel()).to(eng.device)
        sd_d = torch.as_tensor(sd.ravel()).to(eng.device)
    _, _, _, acq = eng.acq_sweep(ACQ_IDS[acq_name], mean_d, sd_d, 1, mu_best=mu_best, xi=xi, alpha=alpha, beta=beta,
                                 want_acq=True)
    return acq.cpu().numpy().reshape(mean.shape)


def confidence_bound(gpmodel, X_full, **kwargs):
    """alpha * mean + beta * sd (acqfunc.py:11-34)."""
    alpha = kwargs.get("alpha", 0)
    beta = kwargs.get("beta", 1)
    mean, sd = gpmodel.predict(X_full, verbose=0)
    acq = _sweep(gpmodel, "cb", mean, sd, alpha=alpha, beta=beta)
    return acq, (mean, sd)


def expected_improvement(gpmodel, X_full, X_sparse, **kwargs):
    """imp * Phi(z) + sd * phi(z) with the incumbent taken as the best predicted mean at the
    already-measured points (acqfunc.py:37-63)."""
    xi = kwargs.get("xi", 0.01)
    mean_sample, _ = gpmodel.predict(X_sparse, verbose=0)
    mean_sample_opt = np.nanmax(mean_sample)
    mean, sd = gpmodel.predict(X_full, verbose=0)
    acq = _sweep(gpmodel, "ei", mean, sd, mu_best=mean_sample_opt, xi=xi)
    return acq, (mean, sd)


def probability_of_improvement(gpmodel, X_full, X_sparse, **kwargs):
    """Phi(z).  The reference takes nanmax over the (mean, sd) TUPLE returned by predict
    (acqfunc.py:86-88); its golden test was generated that way, so the incumbent here is the
    maximum over both arrays as well."""
    xi = kwargs.get("xi", 0.01)
    mean_sample = gpmodel.predict(X_sparse, verbose=0)
    mean_sample_opt = np.nanmax(mean_sample)
    mean, sd = gpmodel.predict(X_full, verbose=0)
    acq = _sweep(gpmodel, "poi", mean, sd, mu_best=mean_sample_opt, xi=xi)
    return acq, (mean, sd)
