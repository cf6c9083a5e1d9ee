"""Training-data sampler of the optimality estimators (reference: utilities.gen_data_ndim, utilities.py:14-59).

The reference samples (Table 1 of the manuscript) an orthonormal eigenbasis, eigenvalues uniform in [-1, 1] and a point x
uniform in [0, 1]^dim, forms Q = V diag(lambda) V^T and solves the dim-dimensional SDP
    min sum(Q o X)  s.t.  lambda_min([[1, x^T], [x, X]]) >= 0,  X_ii <= x_i
with Mosek through cvxpy, one problem at a time (1,000 per ~minute).  Here the sampling is the reference's, call for call
(same legacy numpy RNG stream), and all problems are solved in one batch by the GPU barrier solver
(``sdpcs_sdp_solve``, csrc/sdp_kernels.cuh), so MATLAB-free re-training data can be produced without Mosek.
Row format as written by the reference: [eigvecs^T flattened | eigvals | x | Q upper triangle row-major with the
off-diagonal entries doubled | optimal value].
"""
import numpy as np

from . import _capi


def gen_data_ndim(nb_datapoints, dim, savefile=None, rand_seed=7, device=0):
    from scipy.stats import ortho_group
    np.random.seed(rand_seed)
    iu = np.triu_indices(dim)
    rows = np.empty((nb_datapoints, dim * dim + 2 * dim + dim * (dim + 1) // 2 + 1))
    for r in range(nb_datapoints):
        eigvecs = ortho_group.rvs(dim)                                    # utilities.py:32
        eigvals = np.random.uniform(-1, 1, dim)                           # :34
        Q = np.matmul(np.matmul(eigvecs, np.diag(eigvals)), np.transpose(eigvecs))   # :36
        x = np.random.uniform(0, 1, dim)                                  # :38
        Qt = np.triu(Q, 1) + np.triu(Q, 0)                                # :51
        rows[r, :-1] = np.concatenate([eigvecs.T.flatten(), eigvals, x, Qt[iu]])
    if nb_datapoints:
        eng = _capi.Engine(device)
        x_cols = slice(dim * dim + dim, dim * dim + 2 * dim)
        q_cols = slice(dim * dim + 2 * dim, rows.shape[1] - 1)
        rows[:, -1] = eng.sdp_solve(dim, rows[:, x_cols], rows[:, q_cols])    # sum(Q o X) = sum_{i<=j} Qt_ij X_ij
        eng.close()
    if savefile:
        with open(savefile, "a") as f:
            for line in rows:
                f.write(",".join(str(v) for v in line.tolist()) + "\n")
    return rows
