#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE's own dense oracle.

Run in the build container (needs /root/reference; the GPU box does not have it -- the fixtures travel,
this script does not need to):

    python tests/golden/make_golden.py

What is imported from the reference, by file path (bypassing ``manifold_gp/__init__.py`` which needs gpytorch):
  * ``test/_dense_operators.py``            -- graph_laplacian, matern_precision, matern_labeled_precision,
                                              matern_scaled_precision, matern_noisy_precision   (pure torch)
  * ``manifold_gp/utils/load_dataset.py``   -- get_data, groundtruth_from_samples                (numpy+networkx)
and the reference's own fixture ``manifold_gp/data/dumbbell.msh``.

Everything the reference computes through faiss / torch_sparse / linear_operator (absent here) is NOT
golden: the kNN graph that feeds the dense oracle is built by ``oracle.knn_graph`` (parity unpinned) -- the
same ``idx, val`` then feed both the dense reference and every implementation under test, exactly as the
reference's own test does (test/test_laplacian.py:54-56).

All golden outputs are fp64 (``torch.set_default_dtype(float64)`` so the reference's ``torch.eye`` calls
follow); fp32 implementations are compared against them at 1e-5 relative, fp64 ones at 1e-10.
"""

import importlib.util
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    import oracle

    dense = _load(os.path.join(REF, "test/_dense_operators.py"), "ref_dense_operators")
    lds = _load(os.path.join(REF, "manifold_gp/utils/load_dataset.py"), "ref_load_dataset")

    # ---- dataset: manifold_1D_dataset() (load_dataset.py:10-18) without importlib.resources ----------------
    data = lds.get_data(os.path.join(REF, "manifold_gp/data/dumbbell.msh"), "Nodes", "Elements")
    vertices = data["Nodes"][:, 1:-1]
    edges = data["Elements"][:, -2:].astype(int) - 1
    truth, _ = lds.groundtruth_from_samples(vertices, edges)
    sampled_x = torch.from_numpy(vertices).float()
    sampled_y = torch.from_numpy(truth).float()

    # ---- split: test/test_laplacian.py:20-23 ------------------------------------------------------------------
    num_test = 10
    torch.manual_seed(1337)
    test_idx = torch.zeros(sampled_x.shape[0]).scatter_(0, torch.randperm(sampled_x.shape[0])[:num_test], 1).bool()
    train_x, test_x = sampled_x[~test_idx].contiguous(), sampled_x[test_idx].contiguous()
    train_y, test_y = sampled_y[~test_idx].contiguous(), sampled_y[test_idx].contiguous()
    n = train_x.shape[0]

    np.savez_compressed(os.path.join(HERE, "dumbbell_data.npz"),
                        sampled_x=sampled_x.numpy(), sampled_y=sampled_y.numpy(), test_idx=test_idx.numpy())

    torch.set_default_dtype(torch.float64)
    g = torch.Generator().manual_seed(7)
    V = torch.randn(n, 2, generator=g, dtype=torch.float64)
    V[:, 0] = train_y.double()  # the reference tests multiply by train_y (test_laplacian.py:58)
    mask = torch.zeros(n, dtype=torch.bool)
    mask[torch.randperm(n, generator=g)[:200]] = True  # "labeled" rows for the Schur complement

    for k in (10, 50):  # notebook (cell 10) uses 10; test/test_laplacian.py:37 uses 50
        idx, val32 = oracle.knn_graph(train_x, k)
        val = val32.double()
        ev32, ei = oracle.knn_search(train_x, test_x, k)  # out-of-sample query (riemann_kernel.py:138)
        ev = ev32.double()
        out = {"idx": idx.numpy().astype(np.int32), "val": val32.numpy(), "V": V.numpy(), "mask": mask.numpy(),
               "oos_edge_value": ev32.numpy(), "oos_edge_index": ei.numpy().astype(np.int32)}
        # k=50 reproduces test/test_laplacian.py:34-50 (nu=1, eps=0.5, kappa=0.5); k=10 is the notebooks' setting
        param_sets = ((0.5, 1.3, 2), (0.5, 0.5, 1), (0.05, 0.7, 3)) if k == 10 else ((0.5, 0.5, 1),)
        for eps_v, kappa_v, nu in param_sets:
            tag0 = f"e{eps_v}_k{kappa_v}_nu{nu}"
            for normalization in ("symmetric", "randomwalk"):
                for self_loops in (True, False):
                    tag = f"{tag0}_{normalization}_{'sl' if self_loops else 'nosl'}"
                    eps = torch.tensor([[eps_v]], dtype=torch.float64, requires_grad=True)
                    kappa = torch.tensor([[kappa_v]], dtype=torch.float64, requires_grad=True)
                    L, _, deg_un, _, deg = dense.graph_laplacian(idx, val, eps, n, normalization=normalization,
                                                                 self_loops=self_loops)
                    out[f"{tag}_LV"] = (L @ V).detach().numpy()
                    out[f"{tag}_LtV"] = (L.T @ V).detach().numpy()
                    out[f"{tag}_Ldiag"] = L.diag().detach().numpy()
                    out[f"{tag}_deg_unnorm"] = deg_un.detach().numpy()
                    out[f"{tag}_deg"] = deg.detach().numpy()
                    # d/d eps of sum(L^T v)  (test/_test_functions.py:59-63)
                    (geps,) = torch.autograd.grad((L.T @ V[:, :1]).sum(), eps, retain_graph=True)
                    out[f"{tag}_grad_eps_sumLtv"] = geps.numpy()
                    # Matern precision (dense twin _dense_operators.py:27-33)
                    P = dense.matern_precision(L, nu, kappa, deg if normalization == "randomwalk" else None)
                    out[f"{tag}_PV"] = (P @ V).detach().numpy()
                    # wrappers: the dense twins (_dense_operators.py:52-57).  NB the shipped model multiplies by the
                    # outputscale (riemann_gp.py:35, inverse_scale=False); the dense twin divides: store both.
                    oscale = torch.tensor(1.7, requires_grad=True)
                    noise = torch.tensor(0.02, requires_grad=True)
                    Pdiv = dense.matern_scaled_precision(P, oscale)
                    Pmul = P * oscale
                    out[f"{tag}_PdivV"] = (Pdiv @ V).detach().numpy()
                    out[f"{tag}_PmulV"] = (Pmul @ V).detach().numpy()
                    Pn = dense.matern_noisy_precision(Pdiv, noise)
                    out[f"{tag}_PnoisyV"] = (Pn @ V).detach().numpy()
                    Ps = dense.matern_labeled_precision(P, mask)
                    out[f"{tag}_PschurV"] = (Ps @ V[mask]).detach().numpy()
                    if nu == 2 and self_loops:
                        # precision-form NLL and its gradients (test/_test_functions.py:77-81), dense, fp64
                        y = train_y.double()
                        loss = 0.5 * sum([torch.dot(y, torch.mv(Pn, y)), -torch.logdet(Pn), y.size(-1) * math.log(2 * math.pi)])
                        grads = torch.autograd.grad(loss, [eps, kappa, oscale, noise])
                        out[f"{tag}_nll"] = np.array(loss.item())
                        out[f"{tag}_nll_grads"] = np.array([g_.item() for g_ in grads])
                        out[f"{tag}_logdetP"] = np.array(torch.logdet(P).item())
                        out[f"{tag}_Pinv_V"] = torch.linalg.solve(P, V).detach().numpy()
                    if nu == 1 and not self_loops or nu == 2 and self_loops:
                        # eigenpairs + out-of-sample extension (test/_test_functions.py:107-164)
                        with torch.no_grad():
                            m = 20
                            if normalization == "randomwalk":
                                w, U = torch.linalg.eig(L)
                                w, U = torch.real(w), torch.real(U)
                                w, order = torch.sort(w)
                                U = U[:, order]
                            else:
                                w, U = torch.linalg.eigh(L)
                            w, U = w[:m], U[:, :m]
                            out[f"{tag}_evals"] = w.numpy()
                            rows = torch.arange(ei.shape[0]).repeat_interleave(ei.shape[1])
                            cols = ei.reshape(-1)
                            A_ext_un = torch.sparse_coo_tensor(torch.stack([rows, cols]),
                                                               ev.reshape(-1).div(-4 * eps.square()).exp().squeeze(),
                                                               (ei.shape[0], n)).to_dense()
                            d_ext_un = A_ext_un.sum(dim=1)
                            A_ext = torch.mm(d_ext_un.pow(-1).diag(), torch.mm(A_ext_un, deg_un.pow(-1).diag()))
                            d_ext = A_ext.sum(dim=1)
                            if normalization == "symmetric":
                                ext = torch.mm(d_ext.pow(-0.5).diag(), torch.mm(A_ext, deg.pow(-0.5).diag()))
                            else:
                                ext = torch.mm(d_ext.pow(-1.0).diag(), A_ext)
                            out[f"{tag}_evecs"] = U[:, :6].numpy()  # first 6 modes only (fixture size)
                            out[f"{tag}_ext_evecs"] = torch.mm(ext, U).numpy()
        np.savez_compressed(os.path.join(HERE, f"dumbbell_k{k}.npz"), **out)
        print("wrote", f"dumbbell_k{k}.npz", "M =", idx.shape[1], "keys =", len(out))


if __name__ == "__main__":
    main()
