"""Oracle: mBCG (``linear_cg``), Lanczos tridiagonalisation and stochastic-Lanczos-quadrature log-det.
TEST INFRASTRUCTURE ONLY.

These algorithms live in ``linear_operator`` (third party; un-pinned in the reference's ``setup.py:26-32``,
inferred >= 0.5.0 because ``MaskedLinearOperator`` is used at schur_complement_operator.py:9; absent from
/root/reference and from this image).  They are restated from the published sources
(``linear_operator/utils/linear_cg.py``, ``utils/lanczos.py``, ``utils/stochastic_lq.py``,
``functions/_inv_quad_logdet.py``).  PARITY UNPINNED against the reference -- the
reference's tests never exercise them (``test_solve`` is never called; ``test_ml`` / ``test_eigen`` run with
``max_cholesky=2000 >= N``) -- so they are pinned against dense ``torch.linalg.solve / eigh / logdet`` in
``tests/test_oracle_solvers.py``.

Reference call sites: utils/train_model.py:55,67-68; operators/precision_matern_operator.py:53;
operators/schur_complement_operator.py:28; operators/graph_laplacian_operator.py:133-135.
"""

from __future__ import annotations

import warnings

import torch


def linear_cg(matmul, rhs, n_tridiag=0, tolerance=1.0, eps=1e-10, stop_updating_after=1e-10,
              max_iter=1000, max_tridiag_iter=20, initial_guess=None, terminate_cg_by_size=False,
              return_info=False):
    """Modified batched CG (no preconditioner -- the reference operators define none).

    Defaults mirror ``gpytorch.settings``: cg_tolerance=1, max_cg_iterations=1000,
    max_lanczos_quadrature_iterations=20 (callers in the reference pass cg_tolerance=1e-2).
    """
    is_vector = rhs.dim() == 1
    if is_vector:
        rhs = rhs.unsqueeze(-1)
    if initial_guess is None:
        initial_guess = torch.zeros_like(rhs)
    if max_tridiag_iter > max_iter:
        raise RuntimeError("Getting a tridiagonalization larger than the number of CG iterations run is not possible!")
    num_rows = rhs.size(-2)
    n_iter = min(max_iter, num_rows) if terminate_cg_by_size else max_iter
    n_tridiag_iter = min(max_tridiag_iter, num_rows)

    rhs_norm = rhs.norm(2, dim=-2, keepdim=True)
    rhs_is_zero = rhs_norm.lt(eps)
    rhs_norm = rhs_norm.masked_fill(rhs_is_zero, 1)
    rhs = rhs.div(rhs_norm)

    residual = rhs - matmul(initial_guess)
    result = initial_guess.expand_as(residual).contiguous().clone()
    residual_norm = residual.norm(2, dim=-2, keepdim=True)
    has_converged = torch.lt(residual_norm, stop_updating_after)

    if has_converged.all() and not n_tridiag:
        n_iter = 0
    else:
        curr_conjugate_vec = residual.clone()
        residual_inner_prod = residual.mul(residual).sum(-2, keepdim=True)

    if n_tridiag:
        t_mat = torch.zeros(n_tridiag_iter, n_tridiag_iter, n_tridiag, dtype=rhs.dtype)
        prev_alpha_reciprocal = torch.empty(n_tridiag, dtype=rhs.dtype)
        prev_beta = torch.empty(n_tridiag, dtype=rhs.dtype)
    update_tridiag = True
    last_tridiag_iter = 0
    tolerance_reached = False
    k_done = 0

    for k in range(n_iter):
        mvms = matmul(curr_conjugate_vec)
        # alpha_k = r^T r / p^T A p, with the "safe division" masks of the published code
        alpha = curr_conjugate_vec.mul(mvms).sum(-2, keepdim=True)
        is_zero = alpha.lt(eps)
        alpha = alpha.masked_fill(is_zero, 1)
        alpha = residual_inner_prod / alpha
        alpha = alpha.masked_fill(is_zero, 0)
        alpha = alpha.masked_fill(has_converged, 0)
        residual = residual - alpha * mvms
        result = result + alpha * curr_conjugate_vec
        beta = residual_inner_prod
        residual_inner_prod = residual.mul(residual).sum(-2, keepdim=True)
        is_zero = beta.lt(eps)
        beta = beta.masked_fill(is_zero, 1)
        beta = residual_inner_prod / beta
        beta = beta.masked_fill(is_zero, 0)
        curr_conjugate_vec = curr_conjugate_vec * beta + residual

        residual_norm = residual.norm(2, dim=-2, keepdim=True)
        residual_norm = residual_norm.masked_fill(rhs_is_zero, 0)
        has_converged = torch.lt(residual_norm, stop_updating_after)
        k_done = k + 1

        if (k >= min(10, max_iter - 1) and bool(residual_norm.mean() < tolerance)
                and not (n_tridiag and k < min(n_tridiag_iter, max_iter - 1))):
            tolerance_reached = True
            break

        if n_tridiag and k < n_tridiag_iter and update_tridiag:
            alpha_tridiag = alpha.squeeze(-2)[:n_tridiag].clone()
            beta_tridiag = beta.squeeze(-2)[:n_tridiag].clone()
            a_zero = alpha_tridiag.eq(0)
            alpha_reciprocal = alpha_tridiag.masked_fill(a_zero, 1).reciprocal()
            if k == 0:
                t_mat[k, k] = alpha_reciprocal
            else:
                t_mat[k, k] = alpha_reciprocal + prev_beta * prev_alpha_reciprocal
                off = prev_beta.sqrt() * prev_alpha_reciprocal
                t_mat[k, k - 1] = off
                t_mat[k - 1, k] = off
                if t_mat[k - 1, k].max() < 1e-6:
                    update_tridiag = False
            last_tridiag_iter = k
            prev_alpha_reciprocal = alpha_reciprocal.clone()
            prev_beta = beta_tridiag.clone()

    result = result.mul(rhs_norm)
    if not tolerance_reached and n_iter > 0:
        warnings.warn(
            "CG terminated in {} iterations with average residual norm {} which is larger than the tolerance of {}".format(
                k_done, float(residual_norm.mean()), tolerance), RuntimeWarning)
    if is_vector:
        result = result.squeeze(-1)
    info = {"iterations": k_done, "residual_norm": residual_norm.squeeze(-2).clone(), "converged": tolerance_reached}
    if n_tridiag:
        t_mat = t_mat[: last_tridiag_iter + 1, : last_tridiag_iter + 1]
        t_mat = t_mat.permute(2, 0, 1).contiguous()
        return (result, t_mat, info) if return_info else (result, t_mat)
    return (result, info) if return_info else result


def lanczos_tridiag(matmul, max_iter, n, dtype=torch.float32, init_vec=None, tol=1e-5, generator=None):
    """Lanczos with full re-orthogonalisation (``linear_operator.utils.lanczos.lanczos_tridiag``), one start vector.

    Returns ``q_mat[n, j]`` and ``t_mat[j, j]``.
    """
    if init_vec is None:
        init_vec = torch.randn(n, 1, dtype=dtype, generator=generator)
    init_vec = init_vec.reshape(n, 1).to(dtype)
    num_iter = min(max_iter, n)
    q_mat = torch.zeros(num_iter, n, 1, dtype=dtype)
    t_mat = torch.zeros(num_iter, num_iter, 1, dtype=dtype)

    q0 = init_vec / torch.norm(init_vec, 2, dim=0, keepdim=True)
    q_mat[0].copy_(q0)
    r_vec = matmul(q0)
    alpha_0 = q0.mul(r_vec).sum(0)
    r_vec = r_vec - alpha_0.unsqueeze(0).mul(q0)
    beta_0 = torch.norm(r_vec, 2, dim=0)
    t_mat[0, 0].copy_(alpha_0)
    if num_iter > 1:
        t_mat[0, 1].copy_(beta_0)
        t_mat[1, 0].copy_(beta_0)
        q_mat[1].copy_(r_vec.div(beta_0.unsqueeze(0)))
    k = 0
    for k in range(1, num_iter):
        q_prev = q_mat[k - 1]
        q_curr = q_mat[k]
        beta_prev = t_mat[k, k - 1].unsqueeze(0)
        r_vec = matmul(q_curr) - q_prev.mul(beta_prev)
        alpha_curr = q_curr.mul(r_vec).sum(0, keepdim=True)
        t_mat[k, k].copy_(alpha_curr.squeeze(0))
        if (k + 1) < num_iter:
            r_vec = r_vec - alpha_curr.mul(q_curr)
            correction = r_vec.unsqueeze(0).mul(q_mat[: k + 1]).sum(1, keepdim=True)
            correction = q_mat[: k + 1].mul(correction).sum(0)
            r_vec = r_vec - correction
            r_norm = torch.norm(r_vec, 2, dim=0, keepdim=True)
            r_vec = r_vec / r_norm
            beta_curr = r_norm.squeeze(0)
            t_mat[k, k + 1].copy_(beta_curr)
            t_mat[k + 1, k].copy_(beta_curr)
            inner = q_mat[: k + 1].mul(r_vec.unsqueeze(0)).sum(1)
            could_reorth = False
            for _ in range(10):
                if not torch.sum(inner > tol):
                    could_reorth = True
                    break
                correction = r_vec.unsqueeze(0).mul(q_mat[: k + 1]).sum(1, keepdim=True)
                correction = q_mat[: k + 1].mul(correction).sum(0)
                r_vec = r_vec - correction
                r_norm = torch.norm(r_vec, 2, dim=0, keepdim=True)
                r_vec = r_vec / r_norm
                inner = q_mat[: k + 1].mul(r_vec.unsqueeze(0)).sum(1)
            q_mat[k + 1].copy_(r_vec)
            if torch.sum(beta_curr.abs() > 1e-6) == 0 or not could_reorth:
                break
    num_iter = k + 1
    q = q_mat[:num_iter, :, 0].T.contiguous()
    t = t_mat[:num_iter, :num_iter, 0].contiguous()
    return q, t


def lanczos_tridiag_to_diag(t_mat):
    """``eigh`` of the tridiagonal; negative Ritz values -> 1 and their vectors zeroed (published behaviour)."""
    evals, evecs = torch.linalg.eigh(t_mat)
    mask = evals.ge(0)
    evecs = evecs * mask.to(evecs.dtype).unsqueeze(-2)
    evals = evals.masked_fill(~mask, 1)
    return evals, evecs


def lanczos_diagonalization(matmul, n, num_modes, dtype=torch.float32, init_vec=None, generator=None):
    """``GraphLaplacianOperator.diagonalization('lanczos', num_modes)`` for the symmetric normalisation
    (graph_laplacian_operator.py:132-139): 3*num_modes Lanczos steps, Ritz pairs, ``evals[0] = 0``, truncate."""
    max_iter = 3 * num_modes if 3 * num_modes <= n else n
    q, t = lanczos_tridiag(matmul, max_iter, n, dtype=dtype, init_vec=init_vec, generator=generator)
    evals, v = lanczos_tridiag_to_diag(t)
    evecs = q @ v
    evals = evals.clone()
    evals[0] = 0.0
    if num_modes < n:
        evals, evecs = evals[:num_modes], evecs[:, :num_modes]
    return evals, evecs


def inv_quad_logdet(matmul, n, inv_quad_rhs=None, logdet=True, probes=None, num_trace_samples=10,
                    tolerance=1.0, max_iter=1000, max_tridiag_iter=20, dtype=torch.float32, generator=None,
                    return_info=False):
    """``LinearOperator.inv_quad_logdet`` on the CG/SLQ branch (size > max_cholesky_size).

    log|A| ~= N/P * sum_p sum_i (e1^T v_i)^2 log(theta_i) from the mBCG tridiagonals of P unit-norm Gaussian
    probes; inv_quad = sum(rhs * A^-1 rhs).
    """
    cols = []
    n_probe = 0
    if logdet:
        if probes is None:
            probes = torch.randn(n, num_trace_samples, dtype=dtype, generator=generator)
        probes = probes / probes.norm(2, dim=-2, keepdim=True)
        cols.append(probes)
        n_probe = probes.shape[1]
    if inv_quad_rhs is not None:
        cols.append(inv_quad_rhs if inv_quad_rhs.dim() == 2 else inv_quad_rhs.unsqueeze(-1))
    rhs = torch.cat(cols, dim=-1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = linear_cg(matmul, rhs, n_tridiag=n_probe, tolerance=tolerance, max_iter=max_iter,
                        max_tridiag_iter=max_tridiag_iter, return_info=True)
    if n_probe:
        solves, t_mat, info = out
    else:
        solves, info = out
    inv_quad_term = None
    logdet_term = None
    if inv_quad_rhs is not None:
        r = cols[-1]
        inv_quad_term = (solves[:, n_probe:] * r).sum(-2)
        if inv_quad_term.numel() > 0:
            inv_quad_term = inv_quad_term.sum(-1)
    if logdet:
        evals, evecs = lanczos_tridiag_to_diag(t_mat)
        first = evecs[..., 0, :]
        logdet_term = (n / float(n_probe)) * (first.pow(2) * evals.log()).sum(-1).sum(0)
    if return_info:
        return inv_quad_term, logdet_term, info
    return inv_quad_term, logdet_term
