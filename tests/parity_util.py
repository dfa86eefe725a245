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
    # (a floor of 3 elements: on a 256-element bias vector one ill-conditioned element is already 0.4 %)
    assert loose.sum() <= max(3, 1e-3 * loose.size), f"{what}: {loose.sum()} / {loose.size} elements needed the gradient-interval allowance"
    return int(loose.sum())


def branch_gradients(params, states, actions, y, loss_kind, mask1, mask2, rtol=RTOL):
    """Exact (float64) gradients of the batch-mean loss on the piecewise-linear branch the kernel
    took: relu' masks ``mask1``/``mask2`` [N,B,H] come from the kernel's own dh1/dh2 (non-zero
    pattern).  Two correct fp32 implementations may disagree on relu'(z) when z is within
    round-off of 0; this makes that choice an input instead of loosening the tolerance, and
    asserts the masks differ from the oracle's only at such kinks.  SURVEY.md App. A.8 formulas."""
    w1, b1, w2, b2, w3, b3 = (np.asarray(p, np.float64) for p in params)
    x = np.asarray(states, np.float64)
    n, b, _ = x.shape
    z1 = np.einsum("nbd,ndh->nbh", x, w1) + b1[:, None, :]
    h1 = np.maximum(z1, 0)
    z2 = np.einsum("nbh,nhk->nbk", h1, w2) + b2[:, None, :]
    h2 = np.maximum(z2, 0)
    q = np.einsum("nbh,nha->nba", h2, w3) + b3[:, None, :]
    a = np.asarray(actions, np.int64)
    pred = np.take_along_axis(q, a[..., None], 2)[..., 0]
    e = pred - np.asarray(y, np.float64)
    gcoef = 2 * e / b if loss_kind == "mse" else np.clip(e, -1, 1) / b
    dq = np.zeros_like(q)
    np.put_along_axis(dq, a[..., None], gcoef[..., None], 2)
    m1 = np.asarray(mask1, bool)
    m2 = np.asarray(mask2, bool)
    for z, m, name in ((z1, m1, "layer 1"), (z2, m2, "layer 2")):
        kink = np.abs(z) < 4 * rtol * np.abs(z).max()
        # the kernel may only treat a unit as active away from the kink if it is active (a unit it
        # treats as inactive although z > 0 shows up as a gradient mismatch below, unless dq W3^T == 0)
        assert not (m & (z <= 0) & ~kink).any(), f"{name}: relu mask set where the pre-activation is clearly negative"
    dh2_all = np.einsum("nba,nha->nbh", dq, w3)
    dh2 = dh2_all * m2
    dh1 = np.einsum("nbk,nhk->nbh", dh2, w2) * m1
    grads = [np.einsum("nbd,nbh->ndh", x, dh1), dh1.sum(1), np.einsum("nbh,nbk->nhk", h1, dh2), dh2.sum(1),
             np.einsum("nbh,nba->nha", h2, dq), dq.sum(1)]
    return [g.astype(np.float32) for g in grads]


def relu_boundary(params, states, rtol=RTOL):
    """Units whose pre-activation sits within round-off of the ReLU kink for some sampled row.

    relu'(z) flips there between two correct fp32 implementations (different summation order),
    which changes one whole column of that layer's weight gradient.  Returns (cols1[N] list of
    sets, hit2[N] bools): layer-1 columns to leave out of the W1/b1 comparison, and whether a
    layer-2 unit is affected (then the gradients of every earlier tensor move too and the
    network is skipped for that step).  float64 evaluation of the oracle's weights."""
    w1, b1, w2, b2 = (np.asarray(p, np.float64) for p in params[:4])
    x = np.asarray(states, np.float64)
    z1 = np.einsum("nbd,ndh->nbh", x, w1) + b1[:, None, :]
    z2 = np.einsum("nbh,nhk->nbk", np.maximum(z1, 0), w2) + b2[:, None, :]
    cols1, hit2 = [], []
    for i in range(x.shape[0]):
        near1 = np.abs(z1[i]) < 4 * rtol * np.abs(z1[i]).max()
        cols1.append(set(np.nonzero(near1.any(axis=0))[0].tolist()))
        hit2.append(bool((np.abs(z2[i]) < 4 * rtol * np.abs(z2[i]).max()).any()))
    return cols1, hit2


def drop_cols(got, ref, cols):
    """Copy of ``got`` with the given last-axis columns replaced by the reference's."""
    if not cols:
        return got
    out = np.array(got, copy=True)
    idx = sorted(cols)
    out[..., idx] = np.asarray(ref)[..., idx]
    return out
