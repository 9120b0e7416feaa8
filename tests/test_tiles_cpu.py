"""Tile planner for the tensor-core units kernel: the per-tile GEMM A[units x K] . B[K x (528 | 32)] built from the
plan reproduces the per-term evaluation (tests/flat_eval.py) for both flat layouts."""
import numpy as np
import pytest

from oracle import cavour_oracle as orc
from adrates_b200.curves import OISCurve
from adrates_b200.global_types import InterpTypes
from adrates_b200.synthetic import make_book, flatten_book
from adrates_b200.tiles import plan_tiles, packed_index, node_support_masks, TM, NPACK
from tests.flat_eval import eval_flat
from tests.util_trades import make_calibration_swaps


def sym_tables(d, J, C, pairs):
    """Rows [H_n | C_n | G_nn | G_ab] x (528 packed entries + 32 delta columns), scalings as in k_tables."""
    G, R = J.shape
    g = 1e-4 * J / d[:, None]
    Hf = 1e-8 * (C / d[:, None, None] - J[:, :, None] * J[:, None, :] / (d * d)[:, None, None])
    Cf = 1e-8 * C / d[:, None, None]
    idx = np.array([(j, k) for j in range(32) for k in range(j + 1)])
    pk = lambda M: M[..., idx[:, 0], idx[:, 1]]  # noqa: E731
    z = np.zeros((G, 32))
    rows = [np.hstack([pk(Hf), g]), np.hstack([pk(Cf), g]), np.hstack([pk(g[:, :, None] * g[:, None, :]), z])]
    if len(pairs):
        ga, gb = g[pairs[:, 0]], g[pairs[:, 1]]
        Gab = ga[:, :, None] * gb[:, None, :] + gb[:, :, None] * ga[:, None, :]
        rows.append(np.hstack([pk(Gab), np.zeros((len(pairs), 32))]))
    return np.vstack(rows)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("dedup", [True, False])
def test_tile_gemm_matches_per_term_evaluation(ref_curves, dedup, mode):
    cv = ref_curves["gbp_readme_lzr"]
    vd, swaps = make_calibration_swaps(cv)
    curve = OISCurve(vd, swaps, InterpTypes[cv["interp"]])
    book = make_book(curve, 400, seed=9, max_offset_bd=40)
    flat = flatten_book(book, dedup=dedup)
    check_tile_gemm(flat, cv, book.notional, mode)


@pytest.mark.parametrize("dedup", [True, False])
def test_tile_gemm_on_bond_books(ref_curves, dedup):
    """The tile plan of an array-built bond book (annuity + one-term redemption units) through the same emulation."""
    from adrates_b200.bond_book import BondBook
    from tests.test_bond_book_cpu import BOND_CONVS, _random_bonds
    cv = ref_curves["gbp_readme_lzr"]
    vd, swaps = make_calibration_swaps(cv)
    curve = OISCurve(vd, swaps, InterpTypes[cv["interp"]])
    spec = _random_bonds(curve, 150, np.random.default_rng(4))
    book = BondBook.from_arrays(curve, **spec, **BOND_CONVS["semi_act365"])
    flat = book.flatten(dedup=dedup, tiles=False)
    check_tile_gemm(flat, cv, spec["face_value"])


def check_tile_gemm(flat, cv, notional, mode=0):
    """numpy emulation of the tiled Greeks kernel (coefficient tile x symmetric tables) on `flat` with the plan
    `plan_tiles` gives it, against the per-term evaluation of the same flat arrays."""
    plan = orc.plan_path_b(cv["swap_times"], cv["year_fracs"])
    d, J, C = orc.bootstrap_tables(cv["swap_rates"], plan)
    support = node_support_masks(plan["swap"], plan["prev"], plan["acc"])
    for i in range(len(d)):      # the structural masks are exactly the non-zero pattern of the Jacobian rows
        assert [int(support[i]) >> r & 1 for r in range(32)] == [int(x != 0.0) for x in np.pad(J[i], (0, 32 - J.shape[1]))]
    tp = plan_tiles(flat, len(d), support=support, mode=mode)
    assert len(tp.leftover_units) == 0 and tp.tile_mask.shape == (tp.n_tiles,) and tp.mode == mode
    if mode == 1:
        assert np.all(tp.k_row < len(d)) and np.all((tp.k_coef == 1) | (tp.k_coef == 2)) and len(tp.pairs) == 0
    covered = np.sort(tp.tile_units[tp.tile_units >= 0])
    assert np.array_equal(covered, np.arange(flat.n_units))
    T = sym_tables(d, J, C, tp.pairs.reshape(-1, 2))
    L = np.log(d)
    w = flat.weight.reshape(-1, 2)
    nd = flat.node.reshape(-1, 2)
    u_pv, u_dl, u_gm = np.zeros(flat.n_units), np.zeros((flat.n_units, 32)), np.zeros((flat.n_units, 32, 32))
    for t in range(tp.n_tiles):
        ks, kc, P = tp.tile_kstart[t], tp.tile_kcount[t], tp.tile_npos[t]
        rows, pos, coef = tp.k_row[ks:ks + kc], tp.k_pos[ks:ks + kc], tp.k_coef[ks:ks + kc]
        pos2, coef2 = tp.k_pos2[ks:ks + kc], tp.k_coef2[ks:ks + kc]
        two = coef2 >= 0
        assert np.all((pos2[two] >> 5) == (pos[two] >> 5))     # both contributions in one 32-position chunk
        A = np.zeros((TM, kc))
        units = tp.tile_units[t * TM:(t + 1) * TM]
        for s, u in enumerate(units):
            if u < 0:
                continue
            i0 = flat.unit_offsets[u]
            assert flat.unit_offsets[u + 1] - i0 == P
            i = i0 + pos
            p = flat.amt[i] * np.exp(w[i, 0] * L[nd[i, 0]] + w[i, 1] * L[nd[i, 1]])
            table = np.stack([p, p * w[i, 0], p * w[i, 1], p * w[i, 0] ** 2, p * w[i, 1] ** 2, p * w[i, 0] * w[i, 1]])
            A[s] = table[coef, np.arange(kc)]
            i2 = i0 + np.where(two, pos2, 0)
            p2 = flat.amt[i2] * np.exp(w[i2, 0] * L[nd[i2, 0]] + w[i2, 1] * L[nd[i2, 1]])
            table2 = np.stack([p2, p2 * w[i2, 0], p2 * w[i2, 1], p2 * w[i2, 0] ** 2, p2 * w[i2, 1] ** 2, p2 * w[i2, 0] * w[i2, 1]])
            A[s] += np.where(two, table2[np.where(two, coef2, 0), np.arange(kc)], 0.0)
            u_pv[u] = (flat.amt[i0:i0 + P] * np.exp(w[i0:i0 + P, 0] * L[nd[i0:i0 + P, 0]] + w[i0:i0 + P, 1] * L[nd[i0:i0 + P, 1]])).sum()
        Cm = A @ T[rows]
        if mode == 1:        # MODE_SYRK: the rank-one part sum_i p_i v_i v_i^T per unit, v_i = w0 g_a + w1 g_b, evaluated directly
            g = 1e-4 * J / d[:, None]
            gp = np.pad(g, ((0, 0), (0, 32 - g.shape[1])))
            jj, kk = np.array([(j, k) for j in range(32) for k in range(j + 1)]).T
            for s, u in enumerate(units):
                if u < 0:
                    continue
                i = np.arange(flat.unit_offsets[u], flat.unit_offsets[u + 1])
                p = flat.amt[i] * np.exp(w[i, 0] * L[nd[i, 0]] + w[i, 1] * L[nd[i, 1]])
                v = w[i, 0, None] * gp[nd[i, 0]] + w[i, 1, None] * gp[nd[i, 1]]
                Cm[s, :NPACK] += ((p[:, None] * v).T @ v)[jj, kk]
        # column compaction: everything outside the tile's active pillars is structurally zero
        pos_of = np.argsort(tp.perm)              # the masks are in permuted pillar order
        act = np.array([(int(tp.tile_mask[t]) >> int(pos_of[r])) & 1 for r in range(32)], dtype=bool)
        dead = np.array([not (act[j] and act[k]) for j in range(32) for k in range(j + 1)] + list(~act))
        assert not np.any(T[rows][:, :NPACK + 32][:, dead])
        for s, u in enumerate(units):
            if u < 0:
                continue
            u_dl[u] = Cm[s, NPACK:NPACK + 32]
            for j in range(32):
                for k in range(32):
                    u_gm[u, j, k] = Cm[s, packed_index(j, k)]
    # unit results -> trades, then compare with the per-term evaluation
    K = flat.n_comp
    cw = flat.comp_weight.reshape(-1, K)
    pv, dl, gm = eval_flat(flat, d, J, C)
    for gi in range(flat.n_groups):
        ids = flat.group_units[gi * K:(gi + 1) * K]
        for t in range(flat.group_offsets[gi], flat.group_offsets[gi + 1]):
            row = t if flat.out_index is None else flat.out_index[t]
            v = sum(cw[t, k] * u_pv[ids[k]] for k in range(K))
            dd = sum(cw[t, k] * u_dl[ids[k]] for k in range(K))
            gg = sum(cw[t, k] * u_gm[ids[k]] for k in range(K))
            N = notional[row]
            assert abs(v - pv[row]) <= 1e-12 * N
            assert np.max(np.abs(dd - dl[row])) <= 1e-12 * N * 1e-4 * 50
            assert np.max(np.abs(gg - gm[row])) <= 1e-12 * N * 1e-8 * 2500


def abi_plan_checks(flat, n_nodes):
    """The host-side checks of cav_portfolio_set_tiles (csrc/cav_api.cu), restated: returns the largest number of K rows
    any tile has inside one chunk of 32 term positions (limit 160)."""
    from adrates_b200.tiles import tile_class
    tp, off = flat.tile_plan, flat.unit_offsets
    n_rows = 3 * n_nodes + len(tp.pairs) // 2
    tu = tp.tile_units.reshape(-1, TM)
    assert np.all((tu >= -1) & (tu < flat.n_units)) and int((tu >= 0).sum()) == flat.n_units
    cls_prev, worst = 0, 0
    for t in range(tp.n_tiles):
        ks, kc = int(tp.tile_kstart[t]), int(tp.tile_kcount[t])
        assert ks >= 0 and kc >= 0 and ks + kc <= len(tp.k_row)
        lens = {int(off[u + 1] - off[u]) for u in tu[t] if u >= 0}
        assert len(lens) == 1                                    # units of a tile have the same number of terms
        npos = lens.pop()
        assert npos <= 255
        c = tile_class(tp.tile_mask[t])
        assert c >= cls_prev                                     # tiles ordered by size class
        cls_prev = c
        prev, in_chunk = 0, 0
        for k in range(kc):
            p = int(tp.k_pos[ks + k])
            assert prev <= p < npos                              # K rows ordered by position
            in_chunk = in_chunk + 1 if (p >> 5) == (prev >> 5) else 1
            worst = max(worst, in_chunk)
            prev = p
            if tp.k_coef2[ks + k] >= 0:
                p2 = int(tp.k_pos2[ks + k])
                assert 0 <= p2 < npos and (p2 >> 5) == (p >> 5)  # second contribution in the same chunk
    assert worst <= 160
    assert np.all((tp.k_row >= 0) & (tp.k_row < n_rows) & (tp.k_coef >= 0) & (tp.k_coef <= 5) & (tp.k_coef2 <= 5))
    assert np.all((tp.pairs >= 0) & (tp.pairs < n_nodes)) and sorted(tp.perm.tolist()) == list(range(32))
    return worst


@pytest.mark.parametrize("dedup", [True, False])
def test_plans_pass_the_abi_checks(ref_curves, dedup):
    """Plans of OIS and bond books (annual books reach the 160-row limit exactly) satisfy what the C ABI verifies."""
    from adrates_b200.bond_book import BondBook
    from tests.test_bond_book_cpu import BOND_CONVS, _random_bonds
    cv = ref_curves["gbp_readme_lzr"]
    vd, swaps = make_calibration_swaps(cv)
    curve = OISCurve(vd, swaps, InterpTypes[cv["interp"]])
    G = curve.path_b_plan().n_nodes
    assert abi_plan_checks(flatten_book(make_book(curve, 600, seed=3), dedup=dedup), G) <= 160
    for conv in BOND_CONVS.values():
        spec = _random_bonds(curve, 160, np.random.default_rng(31))
        abi_plan_checks(BondBook.from_arrays(curve, **spec, **conv).flatten(dedup=dedup, tiles=True), G)
