"""Tile planner for the tiled units kernel (k_units_mma).

The Greeks of a unit are a linear combination of curve-only matrices.  With v = w0 g_a + w1 g_b and
H_n = hess(ln d_n), a bracketed term contributes

    p (v v^T + w0 H_a + w1 H_b) = p w0 H_a + p w1 H_b + p w0^2 G_aa + p w1^2 G_bb + p w0 w1 G_ab,
    G_aa = g_a g_a^T,  G_ab = g_a g_b^T + g_b g_a^T                       (a grid snap: p (H_a + G_aa) = p C_a)

so for units that bracket the same node pairs position by position ("same signature"),
gamma[units x 528 packed entries | 32 delta columns] = A[units x K] . B[K x 560] with B rows taken from
per-curve symmetric tables and only the K coefficients A depending on the unit.  That is a dense FP64 GEMM
per tile of TM units, which the kernel runs on the tensor pipe (mma.sync.m8n8k4.f64); each table row is read
once per tile instead of once per unit.  Only the columns of the tile's *active* par-rate pillars are
computed (tile_mask, from the dependency structure of the bootstrap plan); tiles are ordered by size class
(number of compact columns / 32) because the kernel is instantiated per class.

This module groups units by signature, cuts the groups into tiles of TM units and emits, per group, the list
of K rows: (table row id, term position, coefficient kind).  Units whose terms are not single-DF brackets
(6-pair product terms) are left to the generic warp-per-unit kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

TM = 16            # units per tile (one CTA of k_units_mma owns a tile)
NCOL = 576         # padded row length: 528 packed gamma entries + 32 delta columns + 16 zeros
NPACK = 528

# coefficient kinds
COEF_P, COEF_PW0, COEF_PW1, COEF_PW0SQ, COEF_PW1SQ, COEF_PW0W1 = range(6)

# Plan modes.  MODE_TABLE: p (v v^T + w0 H_a + w1 H_b) is expanded into up to five table rows per term (H_a, H_b, G_aa, G_bb,
# G_ab).  MODE_SYRK: only the curve-Hessian part sum_m p w_m H_{n_m} goes through table rows (one or two H rows per term, K
# shrinks from ~5 to ~1 row per term once neighbouring terms share a node); the rank-one part sum_i p_i v_i v_i^T,
# v_i = w0 g_a + w1 g_b, is a per-unit symmetric rank-k update the kernel evaluates on chip from the 67 KB g table
# (k_units_syrk) - no table rows, no pair rows.
MODE_TABLE, MODE_SYRK = 0, 1


@dataclass
class TilePlan:
    n_tiles: int
    tile_size: int
    tile_units: np.ndarray    # i32 [n_tiles*tile_size], -1 = padding
    tile_kstart: np.ndarray   # i32 [n_tiles] first K row of the tile's group
    tile_kcount: np.ndarray   # i32 [n_tiles]
    tile_npos: np.ndarray     # i32 [n_tiles] terms per unit in this tile
    k_row: np.ndarray         # i32 [sum K over groups] table row id
    k_pos: np.ndarray         # i32 term position the coefficient comes from
    k_coef: np.ndarray        # i32 coefficient kind
    k_pos2: np.ndarray        # i32 second contribution to the same table row (same 32-position chunk), -1 = none
    k_coef2: np.ndarray       # i32
    perm: np.ndarray          # i32 [32] pillar permutation: position q of the packed tables holds pillar perm[q]
    pairs: np.ndarray         # i32 [n_pair_rows*2] (a, b) node pairs that need a G_ab row
    tile_mask: np.ndarray     # u32 [n_tiles] bit q set = pillar perm[q] can be non-zero in this tile's Greeks
    n_table_rows: int         # 3*G + n_pair_rows (+1 zero row appended by the library)
    leftover_units: np.ndarray  # i32 units not covered by tiles (generic kernel)
    mode: int = 0             # MODE_TABLE: every product of log-DF gradients is a table row; MODE_SYRK: H rows only (see below)

    @property
    def max_k(self) -> int:
        return int(self.tile_kcount.max()) if self.n_tiles else 0


def row_H(n, G):
    return n


def row_C(n, G):
    return G + n


def row_Gnn(n, G):
    return 2 * G + n


def node_support_masks(node_swap, node_prev, node_acc=None) -> np.ndarray:
    """u32 [G]: bit r set when ln d_n can depend on par rate r.  From the bootstrap recursion
    d_i = (1 - r_s P_p)/(1 + r_s a), P_i = P_p + a d_i (engine.py:2341-2347): supp(d_i) = {s} U supp(P_p),
    supp(P_i) = supp(P_p) U supp(d_i).  The second-order tables H_n, C_n live on supp x supp."""
    G = len(node_swap)
    md = np.zeros(G, dtype=np.uint32)
    mp = np.zeros(G, dtype=np.uint32)
    for i in range(G):
        p = int(node_prev[i])
        base = mp[p] if p >= 0 else np.uint32(0)
        if p < 0 and node_acc is not None and node_acc[i] == 0.0:
            continue          # root node: d = 1/(1 + r*0) = 1, no dependence on any rate
        md[i] = base | np.uint32(1 << int(node_swap[i]))
        mp[i] = base | md[i]
    return md


def tile_class(mask: int) -> int:
    """Size class of a tile (mirrors cav_tile_class in csrc/cav_api.cu)."""
    na = bin(int(mask)).count("1")
    nnt = (na * (na + 3) // 2 + 7) // 8
    return 0 if nnt <= 8 else 1 if nnt <= 16 else 2 if nnt <= 24 else 3 if nnt <= 32 else 4 if nnt <= 48 else 5


def plan_tiles(flat, G: int, min_group: int = 1, support: np.ndarray | None = None, merge: bool = True,
               permute: bool = True, mode: int = MODE_TABLE) -> TilePlan:
    """flat: FlatPortfolio with n_pairs == 2.  G: number of curve nodes.  support: u32 [G] pillar masks of the
    nodes (node_support_masks) or None = every pillar is active everywhere (no column compaction)."""
    if flat.n_pairs != 2:
        raise ValueError("tile planner handles single-DF terms only")
    full = np.uint32(0xFFFFFFFF)
    w = flat.weight.reshape(-1, 2)
    nd = flat.node.reshape(-1, 2).astype(np.int64)
    off = flat.unit_offsets
    kind = np.where((w[:, 0] == 1.0) & (w[:, 1] == 0.0), 0, np.where(w[:, 1] == 0.0, 1, 2)).astype(np.int64)
    a = nd[:, 0]
    b = np.where(kind == 2, nd[:, 1], 0)
    key = (kind << 40) | (a << 20) | b
    groups = {}
    for u in range(flat.n_units):
        groups.setdefault(key[off[u]:off[u + 1]].tobytes(), []).append(u)
    pair_index = {}
    k_row, k_pos, k_coef, k_pos2, k_coef2 = [], [], [], [], []
    t_units, t_kstart, t_kcount, t_npos, t_mask, t_work = [], [], [], [], [], []
    leftover = []

    def emit(open_rows, row, pos, cf):
        """A table row met twice inside one 32-position chunk (consecutive terms bracket a common node) becomes
        ONE K row whose coefficient is the sum of both contributions."""
        hit = open_rows.get(row) if merge else None
        if hit is not None and k_pos2[hit] < 0 and (k_pos[hit] >> 5) == (pos >> 5):
            k_pos2[hit] = pos
            k_coef2[hit] = cf
            return
        open_rows[row] = len(k_row)
        k_row.append(row); k_pos.append(pos); k_coef.append(cf); k_pos2.append(-1); k_coef2.append(-1)

    for sig, units in groups.items():
        if len(units) < min_group:
            leftover += units
            continue
        ks = np.frombuffer(sig, dtype=np.int64)
        kstart = len(k_row)
        mask = full
        if support is not None:
            kn = (ks >> 40)
            na_all = ((ks >> 20) & 0xFFFFF)
            nb_all = (ks & 0xFFFFF)[kn == 2]
            mask = np.bitwise_or.reduce(support[na_all]) if len(na_all) else np.uint32(0)
            if len(nb_all):
                mask = mask | np.bitwise_or.reduce(support[nb_all])
        open_rows = {}
        for j, kk in enumerate(ks):
            knd, na, nb = int(kk >> 40), int((kk >> 20) & 0xFFFFF), int(kk & 0xFFFFF)
            if mode == MODE_SYRK:
                emit(open_rows, row_H(na, G), j, COEF_PW0)          # a grid snap has w0 = 1: p H_a (+ p g_a g_a^T on chip)
                if knd == 2:
                    emit(open_rows, row_H(nb, G), j, COEF_PW1)
            elif knd == 0:
                emit(open_rows, row_C(na, G), j, COEF_P)
            else:
                emit(open_rows, row_H(na, G), j, COEF_PW0)
                emit(open_rows, row_Gnn(na, G), j, COEF_PW0SQ)
                if knd == 2:
                    pi = pair_index.setdefault((na, nb), len(pair_index))
                    emit(open_rows, row_H(nb, G), j, COEF_PW1)
                    emit(open_rows, row_Gnn(nb, G), j, COEF_PW1SQ)
                    emit(open_rows, 3 * G + pi, j, COEF_PW0W1)
        kcount = len(k_row) - kstart
        ua = np.array(units, dtype=np.int32)
        pad = (-len(ua)) % TM
        ua = np.concatenate([ua, np.full(pad, -1, dtype=np.int32)]).reshape(-1, TM)
        for row in ua:
            t_units.append(row)
            t_kstart.append(kstart)
            t_kcount.append(kcount)
            t_npos.append(len(ks))
            t_mask.append(mask)
            t_work.append(kcount)
    n_tiles = len(t_units)
    # Pillar permutation: pillars that are active in most of the work come first, so that the active sets (which
    # are nested: every pillar up to the maturity, plus a few short-end ones) are prefixes of the permuted order
    # and the compact columns of a tile are mostly CONTIGUOUS in the packed tables (coalesced table-row loads).
    perm = np.arange(32, dtype=np.int32)
    if n_tiles and support is not None and permute:
        m = np.array(t_mask, dtype=np.uint64)
        wk = np.array(t_work, dtype=np.float64)
        freq = np.array([float(np.sum(wk[((m >> np.uint64(r)) & np.uint64(1)) == 1])) for r in range(32)])
        perm = np.argsort(-freq, kind="stable").astype(np.int32)
        pos_of = np.empty(32, dtype=np.int64)
        pos_of[perm] = np.arange(32)
        pm = np.zeros(n_tiles, dtype=np.uint64)
        for r in range(32):
            pm |= ((m >> np.uint64(r)) & np.uint64(1)) << np.uint64(pos_of[r])
        t_mask = [np.uint32(x) for x in pm]
    if n_tiles:      # order tiles by size class (stable: neighbours keep sharing table rows)
        order = np.argsort(np.array([tile_class(m) for m in t_mask]), kind="stable")
        t_units = [t_units[i] for i in order]
        t_kstart = [t_kstart[i] for i in order]
        t_kcount = [t_kcount[i] for i in order]
        t_npos = [t_npos[i] for i in order]
        t_mask = [t_mask[i] for i in order]
    pairs = np.zeros((len(pair_index), 2), dtype=np.int32)
    for (na, nb), pi in pair_index.items():
        pairs[pi] = (na, nb)
    i32 = lambda x: np.array(x, dtype=np.int32)  # noqa: E731
    return TilePlan(
        n_tiles, TM,
        np.concatenate(t_units).astype(np.int32) if n_tiles else np.zeros(0, dtype=np.int32),
        i32(t_kstart), i32(t_kcount), i32(t_npos), i32(k_row), i32(k_pos), i32(k_coef), i32(k_pos2), i32(k_coef2), perm,
        pairs.reshape(-1), np.array(t_mask, dtype=np.uint32), 3 * G + len(pair_index),
        np.array(leftover, dtype=np.int32), mode)


def packed_index(j: int, k: int) -> int:
    j, k = (j, k) if j >= k else (k, j)
    return j * (j + 1) // 2 + k
