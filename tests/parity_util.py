"""Tolerance helpers shared by the GPU parity tests."""
import numpy as np
import torch

RTOL = 1e-5   # BASELINE.json: "within 1e-5 relative for fp32 Q-values, losses and post-Adam weights"


def close(got, ref, rtol=RTOL, scale=None, what=""):
    """|got-ref| <= rtol * max(|ref| elementwise, magnitude of the tensor)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    mag = float(np.abs(ref).max()) if scale is None else scale
    tol = rtol * np.maximum(np.abs(ref), mag if mag > 0 else 1.0)
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} off, worst {np.abs(got - ref).max():.3e} (mag {mag:.3e})"


def adam_close(got, theta0, m0, v0, grad, alpha, eps, rtol=RTOL, what=""):
    """Post-Adam weights against the oracle, honouring Adam's conditioning.

    Adam divides by sqrt(v): where a gradient element is within round-off of zero the step
    (|step| <= ~lr) is not determined by the inputs to 1e-5 -- the quotient g/(|g|+eps') is.
    So the oracle update is evaluated at g and g +- delta with delta = rtol * max|g| (what
    "gradients equal to 1e-5 relative" means for this tensor) and the kernel's weight must
    lie in that interval, widened by rtol * max(|theta|, max|theta|)."""
    from oracle.dqn import adam_update_
    g = torch.as_tensor(grad)
    delta = rtol * float(g.abs().max())
    cands = []
    for s in (-1.0, 0.0, 1.0):
        p, m, v = theta0.clone(), m0.clone(), v0.clone()
        adam_update_(p, g + s * delta, m, v, alpha, eps)
        cands.append(p.numpy().astype(np.float64))
    lo, hi = np.minimum.reduce(cands), np.maximum.reduce(cands)
    ref = cands[1]
    got = np.asarray(got, np.float64)
    tol = rtol * np.maximum(np.abs(ref), float(np.abs(ref).max()))
    bad = (got < lo - tol) | (got > hi + tol)
    loose = np.abs(got - ref) > tol
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} outside the Adam interval, worst {np.abs(got - ref).max():.3e}"
    # the ill-conditioned elements must stay a vanishing minority
    assert loose.mean() < 1e-3, f"{what}: {loose.sum()} / {loose.size} elements needed the gradient-interval allowance"
    return int(loose.sum())
