"""TEST-ONLY numpy evaluation of a FlatPortfolio (the arrays handed to the CUDA library).
Lets the CPU test-suite check the host planner/flattener against the oracle without a GPU;
it restates the kernel's per-term formulas (cav_kernels.cuh, k_units) in dense numpy."""
import numpy as np


def eval_flat(flat, d, J, C):
    """d[G], J[G,R], C[G,R,R] (oracle tables) -> per-trade pv[N], delta[N,R], gamma[N,R,R] in output order."""
    G, R = J.shape
    L = np.log(d)
    g = 1e-4 * J / d[:, None]
    Hf = 1e-8 * (C / d[:, None, None] - J[:, :, None] * J[:, None, :] / (d * d)[:, None, None])
    NP = flat.n_pairs
    w = flat.weight.reshape(-1, NP)
    n = flat.node.reshape(-1, NP)
    ell = np.sum(w * L[n], axis=1)
    p = flat.amt * np.exp(ell)
    v = np.einsum("tm,tmr->tr", w, g[n])
    U = flat.n_units
    upv = np.zeros(U)
    udl = np.zeros((U, R))
    ugm = np.zeros((U, R, R))
    for u in range(U):
        s, e = flat.unit_offsets[u], flat.unit_offsets[u + 1]
        upv[u] = p[s:e].sum()
        udl[u] = p[s:e] @ v[s:e]
        ugm[u] = np.einsum("t,tj,tk->jk", p[s:e], v[s:e], v[s:e]) + np.einsum("t,tm,tmjk->jk", p[s:e], w[s:e], Hf[n[s:e]])
    K = flat.n_comp
    cw = flat.comp_weight.reshape(-1, K)
    N = flat.n_trades
    pv = np.zeros(N)
    dl = np.zeros((N, R))
    gm = np.zeros((N, R, R))
    for gi in range(flat.n_groups):
        ids = flat.group_units[gi * K:(gi + 1) * K]
        for t in range(flat.group_offsets[gi], flat.group_offsets[gi + 1]):
            row = t if flat.out_index is None else flat.out_index[t]
            for k in range(K):
                pv[row] += cw[t, k] * upv[ids[k]]
                dl[row] += cw[t, k] * udl[ids[k]]
                gm[row] += cw[t, k] * ugm[ids[k]]
    return pv, dl, gm
