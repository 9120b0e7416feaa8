// cav_kernels.cuh - sm_100a kernels of the valuation-and-Greeks path (FP64 throughout).
//
//   k_bootstrap      engine grid bootstrap + exact 1st/2nd-order tangents
//                    (reference: Engine.build_curve_ad scan + jacrev/hessian, engine.py:2337-2389)
//   k_tables         ln-DF tables the valuation kernels read: L, g = 1e-4 J/d,
//                    Hf = 1e-8 (C/d - J J^T/d^2) = hess(ln d), Cf = 1e-8 C/d
//   k_units<NP,D,G>  fused interpolation + PV + delta ladder + 32x32 gamma per unit, one warp
//                    per unit (reference: _price_fixed_leg_jax/_float_leg_jax + grad/hessian +
//                    chain rule, engine.py:2414-2448, 2639-2728, 2551-2568)
//   k_expand         per-trade outputs = weighted sums of unit outputs (streaming stores)
//   k_reduce_partials  deterministic portfolio totals (reference: Portfolio.compute sums)
//   k_df_ad          DiscountCurve._linear_forward_interp (discount_curve.py:385-415)
//   k_scen_*         scenario re-bootstrap + revaluation (Model.scenario loop, models.py:507-557)
#pragma once
#include "cav_ctx.h"


// ------------------------------------------------------------------------------------------
// Bootstrap with tangents.  One CTA of 1024 threads; thread (j,k) owns Hessian entry [j][k],
// threads with j == 0 also own Jacobian entry [k].  Nodes are sequential (each depends on
// the annuity of an earlier node), the R x R tangent algebra is parallel.
//   u = 1 - r P_p, v = 1 + r a, d = u/v, P = P_p + a d            (engine.py:2341-2347)
//   dd = (du - d dv)/v,  d2d = (d2u - dd dv^T - dv dd^T)/v
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
k_bootstrap(int G, int order, const double* __restrict__ rates, const double* __restrict__ acc,
            const int* __restrict__ swap, const int* __restrict__ prev,
            double* df, double* P, double* jac, double* dP, double* hess, double* d2P)
{
    const int tid = threadIdx.x;
    const int j = tid >> 5, k = tid & 31;
    for (int i = 0; i < G; ++i) {
        const int s = swap[i];
        const int p = prev[i];
        const double r = rates[s];
        const double a = acc[i];
        const double Pp = (p < 0) ? 0.0 : P[p];
        const double u = 1.0 - r * Pp;
        const double v = 1.0 + r * a;
        const double d = u / v;
        if (tid == 0) { df[i] = d; P[i] = Pp + a * d; }
        if (order >= 1) {
            const double dPp_j = (p < 0) ? 0.0 : dP[p * CAV_RW + j];
            const double dPp_k = (p < 0) ? 0.0 : dP[p * CAV_RW + k];
            const double du_j = -((j == s ? Pp : 0.0) + r * dPp_j);
            const double du_k = -((k == s ? Pp : 0.0) + r * dPp_k);
            const double dv_j = (j == s) ? a : 0.0;
            const double dv_k = (k == s) ? a : 0.0;
            const double dd_j = (du_j - d * dv_j) / v;
            const double dd_k = (du_k - d * dv_k) / v;
            if (j == 0) { jac[i * CAV_RW + k] = dd_k; dP[i * CAV_RW + k] = dPp_k + a * dd_k; }
            if (order >= 2) {
                const double d2Pp = (p < 0) ? 0.0 : d2P[(size_t)p * CAV_RR + tid];
                const double d2u = -((j == s ? dPp_k : 0.0) + (k == s ? dPp_j : 0.0) + r * d2Pp);
                const double d2d = (d2u - dd_j * dv_k - dv_j * dd_k) / v;
                hess[(size_t)i * CAV_RR + tid] = d2d;
                d2P[(size_t)i * CAV_RR + tid] = d2Pp + a * d2d;
            }
        }
        __syncthreads();   // node i is visible to the CTA before any later node reads it
    }
}

// ------------------------------------------------------------------------------------------
// The same recursion, entry-parallel.  Entry [j][k] of a node's Hessian needs only entry [j][k] of its annuity node and
// first-order vectors:  d2P_i = (1 - a r / v) d2P_p + (row / column s terms),  hess_i = -(r / v) d2P_p + ...   so nothing ever
// crosses Hessian rows.  One single-warp CTA per Hessian row j (lane = column k) walks the G nodes with its history in SHARED
// memory - the annuity P, its gradient dP[32] and row j of its Hessian d2P[j][32] of every node that a later node refers to
// (`slot`, assigned on the host; <= one per distinct coupon date) - and recomputes the first-order quantities itself (32-fold
// redundant, a few flops).  No CTA barrier, no global round trip per node: k_bootstrap spent 517 us on 264 nodes x
// (__syncthreads + write + re-read through L2).  What is left is the recursion's own dependency chain P_p -> d = u / v -> P_i:
// one FP64 division per node.  The discount factors keep that exact division (they are compared with the reference to
// 4e-16); the tangents multiply by 1 / v, which depends on the plan and the rates only and is computed for all nodes up
// front, off the chain (1 ulp away from a division, far inside the 1e-12 the tables are held to).  Row 0's CTA also writes df,
// P, jac, dP.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_bootstrap_rows(int G, int order, int n_slots, const double* __restrict__ rates, const double* __restrict__ acc,
                 const int* __restrict__ swap, const int* __restrict__ prev, const int* __restrict__ slot,
                 double* df, double* P, double* jac, double* dP, double* hess)
{
    extern __shared__ double sh[];
    double* hP = sh;                                   // [n_slots]
    double* hdP = hP + ((n_slots + 1) & ~1);           // [n_slots][32]
    double* hd2 = hdP + (size_t)n_slots * CAV_RW;      // [n_slots][32]   row j of d2P
    double* s_acc = hd2 + (size_t)n_slots * CAV_RW;    // [G]
    double* s_inv = s_acc + G;                         // [G] 1 / (1 + r a)
    int* s_swap = reinterpret_cast<int*>(s_inv + G);   // [G]
    int* s_prev = s_swap + G;
    int* s_slot = s_prev + G;
    const int j = blockIdx.x, k = threadIdx.x;
    for (int i = k; i < G; i += 32) {
        s_acc[i] = acc[i]; s_swap[i] = swap[i]; s_prev[i] = prev[i]; s_slot[i] = slot[i];
        s_inv[i] = 1.0 / (1.0 + rates[swap[i]] * acc[i]);
    }
    const double my_rate = rates[k];
    __syncwarp();
    for (int i = 0; i < G; ++i) {
        const int s = s_swap[i], p = s_prev[i];
        const int ps = p < 0 ? -1 : s_slot[p];
        const double r = __shfl_sync(0xffffffffu, my_rate, s);
        const double a = s_acc[i];
        const double Pp = (ps < 0) ? 0.0 : hP[ps];
        const double u = 1.0 - r * Pp;
        const double v = 1.0 + r * a;
        const double d = u / v;
        const double Pi = Pp + a * d;
        double dPi_k = 0.0, d2Pi = 0.0;
        if (order >= 1) {
            const double dPp_j = (ps < 0) ? 0.0 : hdP[ps * CAV_RW + j];
            const double dPp_k = (ps < 0) ? 0.0 : hdP[ps * CAV_RW + k];
            const double du_j = -((j == s ? Pp : 0.0) + r * dPp_j);
            const double du_k = -((k == s ? Pp : 0.0) + r * dPp_k);
            const double dv_j = (j == s) ? a : 0.0;
            const double dv_k = (k == s) ? a : 0.0;
            const double iv = s_inv[i];
            const double dd_j = (du_j - d * dv_j) * iv;
            const double dd_k = (du_k - d * dv_k) * iv;
            dPi_k = dPp_k + a * dd_k;
            if (j == 0) { jac[i * CAV_RW + k] = dd_k; dP[i * CAV_RW + k] = dPi_k; }
            if (order >= 2) {
                const double d2Pp = (ps < 0) ? 0.0 : hd2[ps * CAV_RW + k];
                const double d2u = -((j == s ? dPp_k : 0.0) + (k == s ? dPp_j : 0.0) + r * d2Pp);
                const double d2d = (d2u - dd_j * dv_k - dv_j * dd_k) * iv;
                hess[(size_t)i * CAV_RR + j * CAV_RW + k] = d2d;
                d2Pi = d2Pp + a * d2d;
            }
        }
        if (j == 0 && k == 0) { df[i] = d; P[i] = Pi; }
        const int si = s_slot[i];
        __syncwarp();                                  // everyone has read the history before a slot may be rewritten
        if (si >= 0) {
            if (k == 0) hP[si] = Pi;
            hdP[si * CAV_RW + k] = dPi_k;
            hd2[si * CAV_RW + k] = d2Pi;
        }
        __syncwarp();
    }
}

// One CTA per node, thread (j,k).
__global__ void __launch_bounds__(1024)
k_tables(int order, const double* __restrict__ df, const double* __restrict__ jac,
         const double* __restrict__ hess, double* L, double* g, double* Hf, double* Cf)
{
    const int i = blockIdx.x, tid = threadIdx.x;
    const int j = tid >> 5, k = tid & 31;
    const double d = df[i];
    const double inv = 1.0 / d;
    if (tid == 0) L[i] = log(d);
    if (order >= 1) {
        const double Jj = jac[i * CAV_RW + j], Jk = jac[i * CAV_RW + k];
        if (j == 0) g[i * CAV_RW + k] = 1e-4 * (Jk * inv);
        if (order >= 2) {
            const double c = hess[(size_t)i * CAV_RR + tid] * inv;
            Cf[(size_t)i * CAV_RR + tid] = 1e-8 * c;
            Hf[(size_t)i * CAV_RR + tid] = 1e-8 * (c - (Jj * inv) * (Jk * inv));
        }
    }
}

// ------------------------------------------------------------------------------------------
// Units kernel.  One warp per unit, persistent warps (unit u handled by warp u mod #warps, so
// the accumulation order of the portfolio partials is fixed).  Lanes = terms while evaluating
// exp(); lanes = pillars (delta) / gamma columns while accumulating Greeks.
//   ln DF_i = sum_m w_im L[n_im];  p_i = amt_i exp(ln DF_i)
//   grad   += p_i v_i,                      v_i = sum_m w_im g[n_im]
//   hess   += p_i (v_i v_i^T + sum_m w_im Hf[n_im])     (Cf[n] directly for a pure grid snap)
// ------------------------------------------------------------------------------------------
struct UnitsArgs {
    int64_t n_units;
    const int64_t* unit_offsets;
    const double* amt;
    const double* weight;      // [n_terms][NP]
    const int* node;           // [n_terms][NP]
    const double* L;           // [G]
    const double* g;           // [G][32]
    const double* Hf;          // [G][1024]
    const double* Cf;          // [G][1024]
    const double* unit_weight; // [n_units] sum of trade weights on this unit (portfolio totals)
    const int64_t* out_index;  // direct mode: row of unit u in the outputs (or null = u)
    double* out_pv;            // [rows]
    double* out_delta;         // [rows][32]
    double* out_gamma;         // [rows][1024]
    double* partials;          // [unit slots][1057] or null
};

// A unit is handled by 32/ROWS independent warps; each owns ROWS rows of the 32x32 gamma
// (lane = column) and re-derives the unit's term scalars itself (lanes = terms for the exp(),
// then one shuffle-broadcast per term), so warps never synchronise with each other.  Warp
// slots are persistent (slot s takes units s, s+S, ...), which fixes the accumulation order of
// the portfolio partials.  ROWS trades registers (occupancy) against redundant per-term work.
template <int NP, bool DELTA, bool GAMMA, int ROWS>
__global__ void __launch_bounds__(256)
k_units(UnitsArgs A)
{
    constexpr int WPU = CAV_RW / ROWS;            // warps per unit
    constexpr int SPC = 8 / WPU;                  // unit slots per CTA
    // portfolio partials of a slot: PV + ladder (+ gamma).  Without gamma only the first 33 entries exist: the slot tiles
    // shrink from 8.4 KB to 320 bytes of shared memory and the partial rows the totals kernel reads from 80 MB to 2.5 MB
    constexpr int NTOT = GAMMA ? CAV_NOUT : (1 + CAV_RW);
    constexpr int TSTRIDE = GAMMA ? CAV_NOUT : 40;
    __shared__ double vbuf[8][CAV_RW];
    // portfolio partials live in shared memory (one 1057-double tile per unit slot), not in
    // registers: keeping 32 more accumulators per thread starves the batch of table-row loads
    extern __shared__ double s_tot[];             // [SPC][TSTRIDE] when A.partials != null
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * 8 + wib;
    const int64_t slot = gw / WPU;                 // unit slot of this warp
    const int part = (int)(gw % WPU);              // which row block of the unit
    const int r0 = part * ROWS;
    const int64_t n_slots = (int64_t)gridDim.x * 8 / WPU;
    const bool lead = (part == 0);
    double* my_tot = s_tot + (size_t)(wib / WPU) * TSTRIDE;
    if (A.partials) {
        for (int e = threadIdx.x; e < SPC * TSTRIDE; e += 256) s_tot[e] = 0.0;
        __syncthreads();
    }

    for (int64_t u = slot; u < A.n_units; u += n_slots) {
        const int64_t t0 = A.unit_offsets[u], t1 = A.unit_offsets[u + 1];
        double pv = 0.0, delta = 0.0;
        double acc[GAMMA ? ROWS : 1];
        if (GAMMA) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] = 0.0;
        }
        for (int64_t base = t0; base < t1; base += 32) {
            const int64_t i = base + lane;
            const int cnt = (int)((t1 - base) < 32 ? (t1 - base) : 32);
            double w[NP];
            int n[NP];
            double p = 0.0;
            if (i < t1) {
                double ell = 0.0;
#pragma unroll
                for (int m = 0; m < NP; ++m) {
                    w[m] = A.weight[i * NP + m];
                    n[m] = A.node[i * NP + m];
                    ell += w[m] * A.L[n[m]];
                }
                p = A.amt[i] * exp(ell);
            } else {
#pragma unroll
                for (int m = 0; m < NP; ++m) { w[m] = 0.0; n[m] = 0; }
            }
            pv += p;
            if (DELTA || GAMMA) {
                for (int jj = 0; jj < cnt; ++jj) {
                    const double pj = __shfl_sync(0xffffffffu, p, jj);
                    double wj[NP];
                    int nj[NP];
#pragma unroll
                    for (int m = 0; m < NP; ++m) {
                        wj[m] = __shfl_sync(0xffffffffu, w[m], jj);
                        nj[m] = __shfl_sync(0xffffffffu, n[m], jj);
                    }
                    const bool snapped = (NP == 2) && (wj[0] == 1.0) && (wj[1] == 0.0);
                    double v = 0.0;
#pragma unroll
                    for (int m = 0; m < NP; ++m)
                        if (wj[m] != 0.0) v += wj[m] * __ldg(A.g + (size_t)nj[m] * CAV_RW + lane);
                    if (DELTA && lead) delta += pj * v;
                    if (GAMMA) {
                        if (snapped) {
                            const double* C = A.Cf + (size_t)nj[0] * CAV_RR + r0 * CAV_RW + lane;
#pragma unroll
                            for (int r = 0; r < ROWS; ++r) acc[r] += pj * __ldg(C + r * CAV_RW);
                        } else {
                            __syncwarp();
                            vbuf[wib][lane] = v;
                            __syncwarp();
                            const double pvk = pj * v;
                            const double2* vb = reinterpret_cast<const double2*>(vbuf[wib] + r0);
#pragma unroll
                            for (int r = 0; r < ROWS; r += 2) {
                                const double2 vv = vb[r >> 1];
                                acc[r] += pvk * vv.x;
                                acc[r + 1] += pvk * vv.y;
                            }
#pragma unroll
                            for (int m = 0; m < NP; ++m) {
                                if (wj[m] != 0.0) {
                                    const double pw = pj * wj[m];
                                    const double* H = A.Hf + (size_t)nj[m] * CAV_RR + r0 * CAV_RW + lane;
#pragma unroll
                                    for (int r = 0; r < ROWS; ++r) acc[r] += pw * __ldg(H + r * CAV_RW);
                                }
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
        const int64_t row = A.out_index ? A.out_index[u] : u;
        if (lead) {
            if (A.out_pv && lane == 0) A.out_pv[row] = pv;
            if (DELTA && A.out_delta) A.out_delta[row * CAV_RW + lane] = delta;
        }
        if (GAMMA && A.out_gamma) {
            double* o = A.out_gamma + row * CAV_RR + r0 * CAV_RW + lane;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) __stcs(o + r * CAV_RW, acc[r]);
        }
        if (A.partials) {     // each warp touches only its own rows of its slot's tile: no races
            const double W = A.unit_weight ? A.unit_weight[u] : 1.0;
            if (lead) {
                if (lane == 0) my_tot[0] += W * pv;
                if (DELTA) my_tot[1 + lane] += W * delta;
            }
            if (GAMMA) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) my_tot[33 + (r0 + r) * CAV_RW + lane] += W * acc[r];
            }
        }
    }
    if (A.partials) {          // partial rows keep the stride of the full totals; without gamma only 33 entries are written
        __syncthreads();
        double* P = A.partials + (size_t)blockIdx.x * SPC * CAV_NOUT;
        for (int e = threadIdx.x; e < SPC * NTOT; e += 256) {
            const int sl = e / NTOT, k = e - sl * NTOT;
            P[(size_t)sl * CAV_NOUT + k] = s_tot[sl * TSTRIDE + k];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tiled units kernel (k_units_mma).  For units that bracket the same node pairs term by term,
//   [gamma (528 packed lower-triangle entries) | delta (32)] [units x 560] = A[units x K] . B[K x 560]
// where the K rows of B come from per-curve symmetric tables (built by k_sym_tables / k_pair_tables)
//   H_n = hess(ln d_n)|g_n,  C_n = (H_n + g_n g_n^T)|g_n,  G_nn = g_n g_n^T|0,  G_ab = (g_a g_b^T + g_b g_a^T)|0
// and only the coefficients A (p, p w0, p w1, p w0^2, p w1^2, p w0 w1) depend on the unit
// (adrates_b200/tiles.py).  One CTA = one tile of 16 units, FP64 DMMA (mma.sync.m8n8k4): per chunk of 32 term
// positions the CTA evaluates p = amt*DF once per term, builds A in shared memory, then runs the K loop with
// B fragments read straight from the L2-resident tables (each table row is read once per tile, not once per
// unit); the epilogue stages 8 units at a time in shared memory and writes full symmetric 32x32 rows with
// 32-byte stores.
//
// Column compaction: a tile's Greeks are non-zero only on its active pillars (tile_mask, na bits).  The GEMM
// runs over the na(na+1)/2 packed gamma columns of the active pillars plus their na delta columns instead of
// all 528 + 32: compact column c reads table column sCol[c]; n-tiles of 8 compact columns are dealt round-robin
// to the 8 warps.  The kernel is instantiated per size class NT = n-tiles per warp (accumulator registers and
// the staging rows scale with NT, so small classes run three or four CTAs per SM and their phases overlap).
// Pillars are permuted (PillarPerm, chosen by the planner) so that the active sets are mostly prefixes and the
// compact columns mostly contiguous in the tables.
// ------------------------------------------------------------------------------------------
#define GT_NC 576          // table row length: 528 packed gamma entries + 32 delta columns + 16 zeros
#define GT_NPACK 528

__device__ __forceinline__ int gt_packed(int j, int k) { return j >= k ? j * (j + 1) / 2 + k : k * (k + 1) / 2 + j; }

// rows 0..G-1: H_n | g_n ; G..2G-1: C_n | g_n ; 2G..3G-1: g_n g_n^T | 0   (one CTA per row)
// (struct PillarPerm: cav_ctx.h)

__global__ void __launch_bounds__(GT_NC)
k_sym_tables(int G, const double* __restrict__ g, const double* __restrict__ Hf, const double* __restrict__ Cf,
             double* T, PillarPerm pp)
{
    const int row = blockIdx.x, e = threadIdx.x;
    const int type = row / G, n = row % G;
    double v = 0.0;
    if (e < GT_NPACK) {
        int j = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (j * (j + 1) / 2 > e) --j;
        while ((j + 1) * (j + 2) / 2 <= e) ++j;
        const int k = pp.perm[e - j * (j + 1) / 2];
        j = pp.perm[j];
        if (type == 0) v = Hf[(size_t)n * CAV_RR + j * CAV_RW + k];
        else if (type == 1) v = Cf[(size_t)n * CAV_RR + j * CAV_RW + k];
        else v = g[n * CAV_RW + j] * g[n * CAV_RW + k];
    } else if (e < GT_NPACK + CAV_RW) {
        if (type < 2) v = g[n * CAV_RW + pp.perm[e - GT_NPACK]];
    }
    T[(size_t)row * GT_NC + e] = v;
}

// rows 3G + i: g_a g_b^T + g_b g_a^T | 0 for the node pairs the portfolio brackets; last row = zeros
__global__ void __launch_bounds__(GT_NC)
k_pair_tables(int n_pairs, const int* __restrict__ pairs, const double* __restrict__ g, double* Trows, PillarPerm pp)
{
    const int i = blockIdx.x, e = threadIdx.x;
    double v = 0.0;
    if (i < n_pairs && e < GT_NPACK) {
        int j = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (j * (j + 1) / 2 > e) --j;
        while ((j + 1) * (j + 2) / 2 <= e) ++j;
        const int k = pp.perm[e - j * (j + 1) / 2];
        j = pp.perm[j];
        const int a = pairs[2 * i], b = pairs[2 * i + 1];
        v = g[a * CAV_RW + j] * g[b * CAV_RW + k] + g[b * CAV_RW + j] * g[a * CAV_RW + k];
    }
    Trows[(size_t)i * GT_NC + e] = v;
}

// pillar-support bit mask of every table row (bit r: the row has a non-zero in a column that involves pillar r)
__global__ void __launch_bounds__(GT_NC)
k_row_masks(const double* __restrict__ T, unsigned* masks)
{
    __shared__ unsigned s_m;
    const int row = blockIdx.x, e = threadIdx.x;
    if (e == 0) s_m = 0u;
    __syncthreads();
    const double v = T[(size_t)row * GT_NC + e];
    unsigned m = 0u;
    if (v != 0.0) {
        if (e < GT_NPACK) {
            int j = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (j * (j + 1) / 2 > e) --j;
            while ((j + 1) * (j + 2) / 2 <= e) ++j;
            m = (1u << j) | (1u << (e - j * (j + 1) / 2));
        } else if (e < GT_NPACK + CAV_RW) m = 1u << (e - GT_NPACK);
    }
    if (m) atomicOr(&s_m, m);
    __syncthreads();
    if (e == 0) masks[row] = s_m;
}

// every K row of every tile must live on the tile's active pillars; *flag != 0 otherwise
__global__ void k_check_tile_masks(int n_tiles, const int* __restrict__ tile_kstart, const int* __restrict__ tile_kcount,
                                   const unsigned* __restrict__ tile_mask, const int2* __restrict__ k_pack,
                                   const unsigned* __restrict__ row_masks, int* flag)
{
    // warp per tile, lanes stride over its K rows (a thread per tile walked ~100 dependent loads: tens of microseconds
    // in front of the units kernel on every new plan)
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= n_tiles) return;
    if (t > 0 && tile_kstart[t] == tile_kstart[t - 1] && tile_kcount[t] == tile_kcount[t - 1] && tile_mask[t] == tile_mask[t - 1])
        return;                                                                                          // same group
    const unsigned allowed = tile_mask[t];
    const int k0 = tile_kstart[t], kc = tile_kcount[t];
    unsigned bad = 0u;
    for (int k = lane; k < kc; k += 32) bad |= row_masks[k_pack[k0 + k].x] & ~allowed;
    if (__any_sync(0xFFFFFFFFu, bad != 0u) && lane == 0) atomicExch(flag, 1);
}

struct SimtArgs {
    const int* tile_units;     // [n_tiles][GT_TM], -1 = padding
    const int* tile_kstart;
    const int* tile_kcount;
    const int* tile_npos;
    const unsigned* tile_mask; // [n_tiles] active par-rate pillars of the tile
    const int2* k_pack;        // per K row: x = table row; y = pos | kind << 8 | pos2 << 16 | kind2 << 24 (kind2 = 7: none)
    PillarPerm pp;             // the packed tables (and the tile masks) are in permuted pillar order
    const double* T;           // symmetric tables [rows][GT_NC]
    const int64_t* unit_offsets;
    const double* amt;
    const double* weight;      // [n_terms][2]
    const int* node;           // [n_terms][2]
    const double* L;
    const double* unit_weight;
    const int64_t* out_index;
    double* out_pv;
    double* out_delta;
    double* out_gamma;
    double* partials;          // [CTAs of this launch][1057] totals per persistent CTA, or null
    const double* term_p = nullptr; // [n_terms] p = amt * DF computed by k_term_scalars ahead of the tile kernels, or null (computed in place)
    double* out_cgamma = nullptr;   // compact unit gammas [unit][GT_NPACK]: packed triangle over the tile's active pillars (k_expand_c)
    unsigned* out_cmask = nullptr;  // [unit] the active-pillar mask its compact row is laid out by
};

// Term scalars p = amt * DF(t) of every term in one streaming pass (experiment, CAV_TERM_PREPASS=1; off by default): the exp
// chain, the two dependent log-DF gathers and the node loads leave the tile kernels and run here at the HBM rate (40 bytes per
// term).  Same expression as in the tile kernels: bit-identical p (tests/test_gpu_units_ws.py).  Measured slower overall on
// private units (3.38 vs 3.30 ms per 1M): the front warps are not what the mma warps wait for.
__global__ void __launch_bounds__(256)
k_term_scalars(int64_t n_terms, const double* __restrict__ amt, const double* __restrict__ weight, const int* __restrict__ node,
               const double* __restrict__ L, double* __restrict__ p_out)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_terms; i += (int64_t)gridDim.x * 256) {
        const double2 w = __ldg(reinterpret_cast<const double2*>(weight) + i);
        const int2 n = __ldg(reinterpret_cast<const int2*>(node) + i);
        const double la = __ldg(L + n.x), lb = __ldg(L + n.y);
        p_out[i] = __ldg(amt + i) * exp(w.x * la + w.y * lb);
    }
}

#define GM_PC 32                 // term positions per chunk
#define GM_KC 160                // K rows per chunk (at most 5 per position)
#define GM_LDA 164               // row stride of the coefficient tile
#ifndef MMA_UNROLL
#define MMA_UNROLL 2
#endif
constexpr int kMmaUnroll = MMA_UNROLL;   // k-steps (4 table rows each) in flight per warp in the K loop

template <int NT>
struct MmaSmem {                 // shared-memory layout (doubles first, then 8-byte and 4-byte integers)
    static constexpr int NCS = 64 * NT;             // compact columns the class can hold
    static constexpr int LDS_ = NCS + 8;            // staging row stride; slot NCS is always zero
    static constexpr int A = 0;                     // [16][GM_LDA]
    static constexpr int TP = A + GT_TM * GM_LDA;   // [16][32] p, then w0, w1
    static constexpr int STAGE = TP + 3 * GT_TM * GM_PC;   // [8][LDS_]
    static constexpr int PV = STAGE + 8 * LDS_;     // [16]
    static constexpr int W = PV + GT_TM;            // [16]
    static constexpr int TOT = W + GT_TM;           // [1058]
    static constexpr int OFF = TOT + 1058;          // int64 [16]
    static constexpr int OUT = OFF + GT_TM;         // int64 [16]
    static constexpr int INTS = OUT + GT_TM;        // int: row[GM_KC] desc[GM_KC] col[NCS] pil[32] unit[16]
    static constexpr size_t BYTES = (size_t)INTS * 8 + (size_t)(2 * GM_KC + NCS + 32 + GT_TM) * 4;
};

template <int NT, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_units_mma(SimtArgs a, int tile_begin, int tile_end, int zero_row)
{
    using SM = MmaSmem<NT>;
    extern __shared__ double smem[];
    double* sA = smem + SM::A;
    double* sTp = smem + SM::TP;
    double* sTw0 = sTp + GT_TM * GM_PC;
    double* sTw1 = sTw0 + GT_TM * GM_PC;
    double* sStage = smem + SM::STAGE;
    double* sPv = smem + SM::PV;
    double* sW = smem + SM::W;
    double* sTot = smem + SM::TOT;                 // portfolio partials of this CTA (each entry owned by one thread)
    int64_t* sOff = reinterpret_cast<int64_t*>(smem + SM::OFF);
    int64_t* sOut = reinterpret_cast<int64_t*>(smem + SM::OUT);
    int* sRow = reinterpret_cast<int*>(smem + SM::INTS);
    int* sDesc = sRow + GM_KC;
    int* sCol = sDesc + GM_KC;
    int* sPil = sCol + SM::NCS;
    int* sUnit = sPil + 32;

    const int tid = threadIdx.x, lane = tid & 31, ng = tid >> 5;
    const int ar = lane >> 2, ac = lane & 3;
    const int oj = tid >> 3, ok4 = (tid & 7) * 4;        // epilogue: this thread owns gamma entries (oj, ok4..ok4+3)
    if (tid < 8) sStage[tid * SM::LDS_ + SM::NCS] = 0.0;
    double* my_tg = sTot + 33 + oj * CAV_RW + ok4;
    my_tg[0] = my_tg[1] = my_tg[2] = my_tg[3] = 0.0;
    if (tid <= CAV_RW) sTot[tid] = 0.0;

    for (int tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        const int K = a.tile_kcount[tile], ks = a.tile_kstart[tile], P = a.tile_npos[tile];
        const unsigned mask = a.tile_mask[tile];
        const int na = __popc(mask), nc = na * (na + 1) / 2, ncols = nc + na, nnt = (ncols + 7) >> 3;
        const int ntw = nnt > ng ? (nnt - ng + 7) >> 3 : 0;              // n-tiles of this warp: ng, ng+8, ...
        __syncthreads();                                  // previous tile is done with shared memory
        if (tid < GT_TM) {
            const int uid = a.tile_units[tile * GT_TM + tid];
            sUnit[tid] = uid;
            sOff[tid] = uid >= 0 ? a.unit_offsets[uid] : 0;
            sOut[tid] = uid >= 0 ? (a.out_index ? a.out_index[uid] : uid) : 0;
            sW[tid] = (uid >= 0 && a.unit_weight) ? a.unit_weight[uid] : 1.0;
            sPv[tid] = 0.0;
        }
        if (tid >= 32 && tid < 64 && ((mask >> (tid - 32)) & 1u)) sPil[__popc(mask & ((1u << (tid - 32)) - 1u))] = tid - 32;
        __syncthreads();
        for (int cc = tid; cc < nnt * 8; cc += 256) {
            int col = GT_NPACK + CAV_RW;                 // a zero column of the tables
            if (cc < nc) {
                int j = (int)((sqrtf(8.0f * cc + 1.0f) - 1.0f) * 0.5f);
                while (j * (j + 1) / 2 > cc) --j;
                while ((j + 1) * (j + 2) / 2 <= cc) ++j;
                const int pj = sPil[j], pk2 = sPil[cc - j * (j + 1) / 2];
                col = pj * (pj + 1) / 2 + pk2;
            } else if (cc < ncols) col = GT_NPACK + sPil[cc - nc];
            sCol[cc] = col;
        }
        __syncthreads();
        int cn[NT];
        double c[2][NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            cn[n] = (n < ntw) ? sCol[(ng + 8 * n) * 8 + ar] : (GT_NPACK + CAV_RW);
            c[0][n][0] = c[0][n][1] = c[1][n][0] = c[1][n][1] = 0.0;
        }

        int kdone = 0;                                   // K rows are ordered by (first) position
        for (int p0 = 0; p0 < P; p0 += GM_PC) {
            __syncthreads();                             // previous chunk consumed
            // (1) term scalars of this chunk: p = amt * DF, w0, w1 for 16 units x 32 positions
            for (int idx = tid; idx < GT_TM * GM_PC; idx += 256) {
                const int u = idx >> 5, j = idx & 31;
                double p = 0.0, w0 = 0.0, w1 = 0.0;
#ifdef MMA_DIAG_NOSCAL      // diagnostic build: no term loads, no exp
                if (p0 + j < P && sUnit[u] >= 0) { p = 1.0 + idx; w0 = 0.75; w1 = 0.25; }
                if (false) {
#else
                if (p0 + j < P && sUnit[u] >= 0) {
#endif
                    const int64_t i = sOff[u] + p0 + j;
                    const double2 w = __ldg(reinterpret_cast<const double2*>(a.weight) + i);
                    w0 = w.x; w1 = w.y;
                    if (a.term_p) p = __ldg(a.term_p + i);
                    else {
                        const int2 n = __ldg(reinterpret_cast<const int2*>(a.node) + i);
                        p = __ldg(a.amt + i) * exp(w.x * __ldg(a.L + n.x) + w.y * __ldg(a.L + n.y));
                    }
                }
                sTp[idx] = p; sTw0[idx] = w0; sTw1[idx] = w1;
            }
            // (2) K rows of this chunk = the prefix of the remaining rows whose position lies in the chunk
            int2 pk = make_int2(zero_row, 0);
            bool in = false;
            if (tid < GM_KC && kdone + tid < K) {
                pk = __ldg(a.k_pack + ks + kdone + tid);
                in = (pk.y & 0xFF) < p0 + GM_PC;
            }
            const int kc = __syncthreads_count(in);      // also publishes the term scalars
            const int kc4 = (kc + 3) & ~3;
            if (tid < kc4) { sRow[tid] = in ? pk.x : zero_row; sDesc[tid] = in ? pk.y : -1; }
            __syncthreads();
            // (3) unit PVs (fixed butterfly order) and the coefficient tile A: warp w owns units 2w, 2w+1
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
                const int u = 2 * ng + uu;
                double pv = sTp[u * GM_PC + lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
                if (lane == 0) sPv[u] += pv;
                for (int k = lane; k < kc4; k += 32) {
                    const int d = sDesc[k];
                    double v = 0.0;
                    if (d >= 0) {
                        const int cf = (d >> 8) & 0xF, cf2 = (d >> 24) & 0xF;
                        const int t = u * GM_PC + (d & 31);
                        const double x0 = sTw0[t], x1 = sTw1[t];
                        const double f1 = (cf == 1 || cf == 3 || cf == 5) ? x0 : ((cf == 2 || cf == 4) ? x1 : 1.0);
                        const double f2 = cf == 3 ? x0 : ((cf == 4 || cf == 5) ? x1 : 1.0);
                        v = sTp[t] * f1 * f2;
                        if (cf2 != 7) {                  // second term feeding the same table row (same chunk)
                            const int t2 = u * GM_PC + ((d >> 16) & 31);
                            const double y0 = sTw0[t2], y1 = sTw1[t2];
                            const double h1 = (cf2 == 1 || cf2 == 3 || cf2 == 5) ? y0 : ((cf2 == 2 || cf2 == 4) ? y1 : 1.0);
                            const double h2 = cf2 == 3 ? y0 : ((cf2 == 4 || cf2 == 5) ? y1 : 1.0);
                            v = fma(sTp[t2] * h1, h2, v);
                        }
                    }
                    sA[u * GM_LDA + k] = v;
                }
            }
            __syncthreads();
            // (4) C += A . B on the tensor pipe
#ifdef MMA_DIAG_NOMMA       // diagnostic build: no K loop
            if (ntw > 0 && a.tile_units == nullptr) {
#else
            if (ntw > 0) {
#endif
                const double* A0 = sA + ar * GM_LDA + ac;
                const double* A1 = A0 + 8 * GM_LDA;
#pragma unroll kMmaUnroll
                for (int k = 0; k < kc4; k += 4) {
                    const double a0 = A0[k], a1 = A1[k];
#ifdef MMA_DIAG_BHOT        // diagnostic build: every B operand from one L1-resident table row
                    const double* rowp = a.T + (size_t)(sRow[k + ac] & 3) * GT_NC;
#else
                    const double* rowp = a.T + (size_t)sRow[k + ac] * GT_NC;
#endif
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        if (n < ntw) {                    // warp-uniform
                            const double b = __ldg(rowp + cn[n]);
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(c[0][n][0]), "+d"(c[0][n][1]) : "d"(a0), "d"(b));
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(c[1][n][0]), "+d"(c[1][n][1]) : "d"(a1), "d"(b));
                        }
                    }
                }
            }
            kdone += kc;
        }
        // ---- epilogue: stage 8 units at a time, scatter to full symmetric rows ----
        int pk4[4], pd = SM::NCS;
        {   // real pillar r lives at position pos_of[r] of the permuted order the mask and the tables use
            const int qj = a.pp.pos_of[oj];
            const int ij = ((mask >> qj) & 1u) ? __popc(mask & ((1u << qj) - 1u)) : -1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int qk = a.pp.pos_of[ok4 + q];
                const int ik = ((mask >> qk) & 1u) ? __popc(mask & ((1u << qk) - 1u)) : -1;
                pk4[q] = (ij < 0 || ik < 0) ? SM::NCS : (ij >= ik ? ij * (ij + 1) / 2 + ik : ik * (ik + 1) / 2 + ij);
            }
            if (tid < CAV_RW) {
                const int qd = a.pp.pos_of[tid];
                if ((mask >> qd) & 1u) pd = nc + __popc(mask & ((1u << qd) - 1u));
            }
        }
        // portfolio partials of this tile accumulate in registers and touch shared memory once per tile: as a
        // read-modify-write per unit they were half of the kernel's shared-memory wavefronts (ncu source counters)
        double tg0 = 0.0, tg1 = 0.0, tg2 = 0.0, tg3 = 0.0, tdl = 0.0, tpv = 0.0;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            __syncthreads();                             // stage rows free
            double* st = sStage + (size_t)ar * SM::LDS_ + 2 * ac;
#pragma unroll
            for (int n = 0; n < NT; ++n)
                if (n < ntw)
                    *reinterpret_cast<double2*>(st + (ng + 8 * n) * 8) = make_double2(c[mt][n][0], c[mt][n][1]);
            __syncthreads();
            for (int s8 = 0; s8 < 8; ++s8) {
                const int u = mt * 8 + s8;
                const int uid = sUnit[u];
                if (uid < 0) continue;
                const double* row_s = sStage + (size_t)s8 * SM::LDS_;
                const int64_t row = sOut[u];
                const double W = sW[u];
                const double g0 = row_s[pk4[0]], g1 = row_s[pk4[1]], g2 = row_s[pk4[2]], g3 = row_s[pk4[3]];
                if (a.out_cgamma) {                       // compact unit row: the staged packed triangle as it is (coalesced)
                    double* cdst = a.out_cgamma + (size_t)uid * GT_NPACK;
                    for (int cc = tid; cc < nc; cc += 256) cdst[cc] = row_s[cc];
                    if (tid == 0) a.out_cmask[uid] = mask;
                }
#ifdef MMA_DIAG_NOSTORE     // diagnostic build: gamma rows gathered from the staging tile but not written
                if (a.out_gamma && g0 == 1.2345e-300) {
#else
                if (a.out_gamma) {
#endif
                    double* dst = a.out_gamma + (size_t)row * CAV_RR + oj * CAV_RW + ok4;
                    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(dst), "d"(g0), "d"(g1), "d"(g2), "d"(g3) : "memory");
                }
                if (a.partials) { tg0 = fma(W, g0, tg0); tg1 = fma(W, g1, tg1); tg2 = fma(W, g2, tg2); tg3 = fma(W, g3, tg3); }
                if (tid < CAV_RW) {
                    const double dl = row_s[pd];
                    if (a.out_delta) a.out_delta[(size_t)row * CAV_RW + tid] = dl;
                    tdl = fma(W, dl, tdl);
                }
                if (tid == 32) {
                    if (a.out_pv) a.out_pv[row] = sPv[u];
                    tpv = fma(W, sPv[u], tpv);
                }
            }
        }
        if (a.partials) {
            my_tg[0] += tg0; my_tg[1] += tg1; my_tg[2] += tg2; my_tg[3] += tg3;
            if (tid < CAV_RW) sTot[1 + tid] += tdl;
            if (tid == 32) sTot[0] += tpv;
        }
    }
    if (a.partials) {     // this CTA's partial row (every launch owns its own block of rows and overwrites it)
        double* Pr = a.partials + (size_t)blockIdx.x * CAV_NOUT;
#pragma unroll
        for (int q = 0; q < 4; ++q) Pr[33 + oj * CAV_RW + ok4 + q] = my_tg[q];
        if (tid < CAV_RW) Pr[1 + tid] = sTot[1 + tid];
        if (tid == 32) Pr[0] = sTot[0];
    }
}

// ------------------------------------------------------------------------------------------
// k_units_mma_ws: the same tile GEMM with warp-specialised CTAs.  In k_units_mma every warp walks the phases of a tile one
// after the other (term scalars, K rows, coefficient tile, K loop, epilogue) and the phases add up: diagnostic builds on
// the private layout give skeleton 1.36 + scalars 0.35 + K loop 1.1 + row stores 0.65 = 3.5 ms per 1M units with the FP64
// tensor pipe idle outside the K loop (profiles/r02_diag_units_mma_phases.txt).  Here warps 8-11 of the CTA ("front")
// prepare chunk c + 1 - term scalars, K rows, coefficient tile A, unit PVs - into the other half of a double buffer
// while warps 0-7 ("mma") run the K loop of chunk c and, after a tile's last chunk, its epilogue.  The halves change
// hands through mbarriers (full / empty); the front warps and the mma warps use named barriers among themselves, there is no
// CTA-wide barrier after the start.  Arithmetic and its order are those of k_units_mma: per-trade results are bit-identical.
// ------------------------------------------------------------------------------------------
#define GW_FRONT 128             // threads of the front warps
#define GW_THREADS (256 + GW_FRONT)
#ifndef GW_NBUF
#define GW_NBUF 3                // coefficient-tile buffers in flight between the front and the mma warps
#endif
#ifndef GW_NSTORE
#define GW_NSTORE 0              // 1 KB row segments per mma warp in flight to HBM as bulk stores (cp.async.bulk); 0 = plain 32-byte
#endif                           // stores.  Measured with 3 slots: 3.65 vs 3.27 ms per 1M private units - the warps wait for the slots.

template <int NT>
struct MmaSmemWs {
    static constexpr int NCS = 64 * NT;
    static constexpr int LDS_ = NCS + 8;
    static constexpr int A = 0;                                  // [GW_NBUF][16][GM_LDA] coefficient tiles
    static constexpr int TP = A + GW_NBUF * GT_TM * GM_LDA;            // front scratch: p, w0, w1 [3][16][32]
    static constexpr int STAGE = TP + 3 * GT_TM * GM_PC;         // mma: [8][LDS_]
    static constexpr int TOT = STAGE + 8 * LDS_;                 // mma: [1058]
    static constexpr int RB = TOT + 1058;                        // mma: [8 warps][GW_NSTORE][128] row segments on their way out
    static constexpr int TW = RB + 8 * GW_NSTORE * 128;          // per tile slot: unit weights [GW_NBUF][16]
    static constexpr int TPV = TW + GW_NBUF * GT_TM;             // per tile slot: unit PVs [GW_NBUF][16]
    static constexpr int PVACC = TPV + GW_NBUF * GT_TM;                // front scratch [16]
    static constexpr int OFF = PVACC + GT_TM;                    // front scratch int64 [16]
    static constexpr int OUT = OFF + GT_TM;                      // per tile slot int64 [GW_NBUF][16]
    static constexpr int BAR = OUT + GW_NBUF * GT_TM;            // mbarriers: full[GW_NBUF], empty[GW_NBUF]
    static constexpr int INTS = BAR + 2 * GW_NBUF;                         // ints from here
    // int layout: row[NBUF][GM_KC] | desc[GM_KC] | col[NBUF][NCS] | pil[32] | unit[NBUF][16] | hdr[NBUF][4] (kc4, last) | tmask[NBUF]
    static constexpr int I_ROW = 0, I_DESC = GW_NBUF * GM_KC, I_COL = I_DESC + GM_KC, I_PIL = I_COL + GW_NBUF * NCS, I_UNIT = I_PIL + 32,
                         I_HDR = I_UNIT + GW_NBUF * GT_TM, I_MASK = I_HDR + 4 * GW_NBUF, I_END = I_MASK + GW_NBUF + (GW_NBUF & 1);
    static constexpr size_t BYTES = (size_t)INTS * 8 + (size_t)I_END * 4;
};

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile("{ .reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=: }" :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bar_named(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ int bar_named_count(int id, int n, bool pred) {
    int c;
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %1, 0; bar.red.popc.u32 %0, %2, %3, p; }" : "=r"(c) : "r"((int)pred), "r"(id), "r"(n) : "memory");
    return c;
}

#ifdef MMA_DIAG_CLOCKS           // diagnostic build: cycles per section, summed per CTA into a.partials-free scratch (g_ws_clocks)
__device__ unsigned long long g_ws_clocks[16];
#define WS_T0() long long t_prev = clock64()
#define WS_T(slot) do { const long long t_now = clock64(); t_acc[slot] += t_now - t_prev; t_prev = t_now; } while (0)
#define WS_TDECL() long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define WS_TFLUSH(base, lead) do { if (lead) for (int q = 0; q < 8; ++q) atomicAdd(&g_ws_clocks[(base) + q], (unsigned long long)t_acc[q]); } while (0)
#else
#define WS_T0()
#define WS_T(slot)
#define WS_TDECL()
#define WS_TFLUSH(base, lead)
#endif

template <int NT, int MINB>
__global__ void __launch_bounds__(GW_THREADS, MINB)
k_units_mma_ws(SimtArgs a, int tile_begin, int tile_end, int zero_row)
{
    using SM = MmaSmemWs<NT>;
    extern __shared__ double smem[];
    double* sAall = smem + SM::A;
    double* sTp = smem + SM::TP;
    double* sTw0 = sTp + GT_TM * GM_PC;
    double* sTw1 = sTw0 + GT_TM * GM_PC;
    double* sStage = smem + SM::STAGE;
    double* sTot = smem + SM::TOT;
    double* sRowBuf = smem + SM::RB;
    double* sWall = smem + SM::TW;
    double* sPvAll = smem + SM::TPV;
    double* sPvAcc = smem + SM::PVACC;
    int64_t* sOff = reinterpret_cast<int64_t*>(smem + SM::OFF);
    int64_t* sOutAll = reinterpret_cast<int64_t*>(smem + SM::OUT);
    uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + SM::BAR);          // full[GW_NBUF], empty[GW_NBUF]
    int* sInt = reinterpret_cast<int*>(smem + SM::INTS);
    int* sRowAll = sInt + SM::I_ROW;
    int* sDesc = sInt + SM::I_DESC;
    int* sColAll = sInt + SM::I_COL;
    int* sPil = sInt + SM::I_PIL;
    int* sUnitAll = sInt + SM::I_UNIT;
    int* sHdr = sInt + SM::I_HDR;
    unsigned* sMask = reinterpret_cast<unsigned*>(sInt + SM::I_MASK);

    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        for (int q = 0; q < GW_NBUF; ++q) { mbar_init(&sBar[q], GW_FRONT / 32); mbar_init(&sBar[GW_NBUF + q], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= 256) {
        // ============================== front warps: up to GW_NBUF chunks ahead of the mma warps ==============================
        // The front is a chain of dependent global loads (tile header -> unit ids -> term offsets -> terms -> log-DFs); what
        // can be fetched ahead is: the next tile's header and unit data travel in registers while the current tile is built,
        // and a chunk's term loads are issued together before the first exp.
        const int ft = tid - 256, fw = ft >> 5;
        WS_TDECL();
        int g = 0;                                               // chunk counter of this CTA (buffer g % NBUF, use g / NBUF)
        int tslot = 0, bslot = 0, buse = 0;                      // tile-data slot; chunk buffer and how often it has been used
        int tile = tile_begin + blockIdx.x;
        int nK = 0, nks = 0, nP = 0; unsigned nmask = 0u;        // header of the next tile
        unsigned last_mask = 0u;                                 // mask of the previous tile (0 = none: an empty mask has no columns)
        int nx_uid = -1; int64_t nx_off = 0, nx_out = 0; double nx_w = 1.0;
        if (tile < tile_end) {
            nK = a.tile_kcount[tile]; nks = a.tile_kstart[tile]; nP = a.tile_npos[tile]; nmask = a.tile_mask[tile];
            if (ft < GT_TM) {
                nx_uid = a.tile_units[tile * GT_TM + ft];
                nx_off = nx_uid >= 0 ? a.unit_offsets[nx_uid] : 0;
                nx_out = nx_uid >= 0 ? (a.out_index ? a.out_index[nx_uid] : nx_uid) : 0;
                nx_w = (nx_uid >= 0 && a.unit_weight) ? a.unit_weight[nx_uid] : 1.0;
            }
        }
        for (; tile < tile_end; tile += gridDim.x) {
            const int K = nK, ks = nks, P = nP;
            const unsigned mask = nmask;
            const int na = __popc(mask), nc = na * (na + 1) / 2, ncols = nc + na, nnt = (ncols + 7) >> 3;
            const int tp = tslot;
            tslot = tslot + 1 == GW_NBUF ? 0 : tslot + 1;
            int* sUnit = sUnitAll + tp * GT_TM;
            int* sCol = sColAll + tp * SM::NCS;
            int kdone = 0;
            const int nchunks = P > 0 ? (P + GM_PC - 1) / GM_PC : 1;
            const int tnext = tile + gridDim.x;
            int nn_uid = -1;
            const unsigned prev_mask = last_mask;
            last_mask = mask;
            for (int ch = 0; ch < nchunks; ++ch, ++g) {
                const int b = bslot, p0 = ch * GM_PC;
                double* sA = sAall + b * GT_TM * GM_LDA;
                int* sRow = sRowAll + b * GM_KC;
                WS_T0();
                mbar_wait(&sBar[GW_NBUF + b], (buse & 1) ^ 1);
                WS_T(0);   // the mma warps are done with this buffer (and, NBUF chunks back, with the tile slot)
                bar_named(2, GW_FRONT);                          // the scratch arrays of the previous chunk are free
                if (ch == 0) {
                    if (ft < GT_TM) {
                        sUnit[ft] = nx_uid; sOff[ft] = nx_off; sOutAll[tp * GT_TM + ft] = nx_out; sWall[tp * GT_TM + ft] = nx_w;
                        sPvAcc[ft] = 0.0;
                        if (tnext < tile_end) nn_uid = a.tile_units[tnext * GT_TM + ft];       // consumed after this chunk's work
                    }
                    if (ft >= 32 && ft < 64 && ((mask >> (ft - 32)) & 1u)) sPil[__popc(mask & ((1u << (ft - 32)) - 1u))] = ft - 32;
                    if (ft == 64) sMask[tp] = mask;
                    if (tnext < tile_end) { nK = a.tile_kcount[tnext]; nks = a.tile_kstart[tnext]; nP = a.tile_npos[tnext]; nmask = a.tile_mask[tnext]; }
                    bar_named(2, GW_FRONT);
                    if (mask == prev_mask) {                     // tiles of a signature group follow each other: same column map
                        const int* pc = sColAll + (tp == 0 ? GW_NBUF - 1 : tp - 1) * SM::NCS;
                        for (int cc = ft; cc < nnt * 8; cc += GW_FRONT) sCol[cc] = pc[cc];
                    } else
                    for (int cc = ft; cc < nnt * 8; cc += GW_FRONT) {
                        int col = GT_NPACK + CAV_RW;             // a zero column of the tables
                        if (cc < nc) {
                            int j = (int)((sqrtf(8.0f * cc + 1.0f) - 1.0f) * 0.5f);
                            while (j * (j + 1) / 2 > cc) --j;
                            while ((j + 1) * (j + 2) / 2 <= cc) ++j;
                            const int pj = sPil[j], pk2 = sPil[cc - j * (j + 1) / 2];
                            col = pj * (pj + 1) / 2 + pk2;
                        } else if (cc < ncols) col = GT_NPACK + sPil[cc - nc];
                        sCol[cc] = col;
                    }
                }
                WS_T(1);
                // (1) term scalars of this chunk: p = amt * DF, w0, w1 for 16 units x 32 positions; (2) K rows of this chunk =
                // the prefix of the remaining rows whose position lies in the chunk.  All first-level loads go out together.
                constexpr int SQ = GT_TM * GM_PC / GW_FRONT;     // scalar slots per front thread
                double2 wv[SQ]; int2 nv[SQ]; double av[SQ]; bool okq[SQ];
#pragma unroll
                for (int q = 0; q < SQ; ++q) {
                    const int idx = ft + q * GW_FRONT, u = idx >> 5, j = idx & 31;
                    okq[q] = p0 + j < P && sUnit[u] >= 0;
                    wv[q] = make_double2(0.0, 0.0); nv[q] = make_int2(0, 0); av[q] = 0.0;
#ifndef MMA_DIAG_NOSCAL
                    if (okq[q]) {
                        const int64_t i = sOff[u] + p0 + j;
                        wv[q] = __ldg(reinterpret_cast<const double2*>(a.weight) + i);
                        if (a.term_p) av[q] = __ldg(a.term_p + i);               // p itself (k_term_scalars)
                        else { nv[q] = __ldg(reinterpret_cast<const int2*>(a.node) + i); av[q] = __ldg(a.amt + i); }
                    }
#endif
                }
                int2 pk0 = make_int2(zero_row, 0), pk1 = make_int2(zero_row, 0);
                bool in0 = false, in1 = false;
                if (kdone + ft < K) { pk0 = __ldg(a.k_pack + ks + kdone + ft); in0 = (pk0.y & 0xFF) < p0 + GM_PC; }
                if (ft + GW_FRONT < GM_KC && kdone + ft + GW_FRONT < K) {
                    pk1 = __ldg(a.k_pack + ks + kdone + ft + GW_FRONT);
                    in1 = (pk1.y & 0xFF) < p0 + GM_PC;
                }
                double la[SQ], lb[SQ];
#pragma unroll
                for (int q = 0; q < SQ; ++q) {
                    la[q] = lb[q] = 0.0;
                    if (!a.term_p) { la[q] = __ldg(a.L + nv[q].x); lb[q] = __ldg(a.L + nv[q].y); }
                }
#pragma unroll
                for (int q = 0; q < SQ; ++q) {
                    const int idx = ft + q * GW_FRONT;
                    double p = 0.0;
#ifdef MMA_DIAG_NOSCAL
                    if (okq[q]) { p = 1.0 + idx; wv[q] = make_double2(0.75, 0.25); }
#else
                    if (okq[q]) p = a.term_p ? av[q] : av[q] * exp(wv[q].x * la[q] + wv[q].y * lb[q]);
#endif
                    sTp[idx] = p; sTw0[idx] = okq[q] ? wv[q].x : 0.0; sTw1[idx] = okq[q] ? wv[q].y : 0.0;
                }
                if (ch == 0 && ft < GT_TM) {                     // second level of the next tile's unit data (its ids have arrived by now)
                    nx_uid = nn_uid;
                    nx_off = nx_uid >= 0 ? a.unit_offsets[nx_uid] : 0;
                    nx_out = nx_uid >= 0 ? (a.out_index ? a.out_index[nx_uid] : nx_uid) : 0;
                    nx_w = (nx_uid >= 0 && a.unit_weight) ? a.unit_weight[nx_uid] : 1.0;
                }
                WS_T(2);
                const int kc = bar_named_count(2, GW_FRONT, in0) + bar_named_count(2, GW_FRONT, in1);   // also publishes the scalars
                const int kc4 = (kc + 3) & ~3;
                if (ft < kc4) { sRow[ft] = in0 ? pk0.x : zero_row; sDesc[ft] = in0 ? pk0.y : -1; }
                if (ft + GW_FRONT < kc4) { sRow[ft + GW_FRONT] = in1 ? pk1.x : zero_row; sDesc[ft + GW_FRONT] = in1 ? pk1.y : -1; }
                bar_named(2, GW_FRONT);
                WS_T(3);
                // (3) unit PVs (fixed butterfly order) and the coefficient tile A: front warp w owns units 4w .. 4w+3.
                // Every FP64 instruction of a front warp queues behind the DMMAs of the mma warps on the same pipe (hundreds of
                // cycles each under load, measured), so dependent FP64 chains are what the front must avoid: a K row's
                // descriptor is decoded once and its coefficient evaluated for the warp's four units side by side.
                {
                    double pv4[4];
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) pv4[uu] = sTp[(4 * fw + uu) * GM_PC + lane];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int uu = 0; uu < 4; ++uu) pv4[uu] += __shfl_xor_sync(0xffffffffu, pv4[uu], o);
                    }
                    if (lane < 4) {
                        const int u = 4 * fw + lane;
                        const double pvl = lane == 0 ? pv4[0] : lane == 1 ? pv4[1] : lane == 2 ? pv4[2] : pv4[3];
                        const double acc = sPvAcc[u] + pvl;
                        sPvAcc[u] = acc;
                        if (ch == nchunks - 1) sPvAll[tp * GT_TM + u] = acc;
                    }
                }
                for (int k = lane; k < kc4; k += 32) {
                    const int d = sDesc[k];
                    const int cf = (d >> 8) & 0xF, cf2 = (d >> 24) & 0xF;
                    const int s1 = (cf == 1 || cf == 3 || cf == 5) ? 0 : ((cf == 2 || cf == 4) ? 1 : 2);       // factor 1: w0 / w1 / 1
                    const int s2 = cf == 3 ? 0 : ((cf == 4 || cf == 5) ? 1 : 2);
                    const int r1 = (cf2 == 1 || cf2 == 3 || cf2 == 5) ? 0 : ((cf2 == 2 || cf2 == 4) ? 1 : 2);
                    const int r2 = cf2 == 3 ? 0 : ((cf2 == 4 || cf2 == 5) ? 1 : 2);
                    const bool two = d >= 0 && cf2 != 7;
                    double v[4];
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        const int t = (4 * fw + uu) * GM_PC + (d & 31);
                        const double x0 = sTw0[t], x1 = sTw1[t];
                        const double f1 = s1 == 0 ? x0 : (s1 == 1 ? x1 : 1.0);
                        const double f2 = s2 == 0 ? x0 : (s2 == 1 ? x1 : 1.0);
                        v[uu] = d >= 0 ? sTp[t] * f1 * f2 : 0.0;
                    }
                    if (two) {                                   // second term feeding the same table row (same chunk)
#pragma unroll
                        for (int uu = 0; uu < 4; ++uu) {
                            const int t2 = (4 * fw + uu) * GM_PC + ((d >> 16) & 31);
                            const double y0 = sTw0[t2], y1 = sTw1[t2];
                            const double h1 = r1 == 0 ? y0 : (r1 == 1 ? y1 : 1.0);
                            const double h2 = r2 == 0 ? y0 : (r2 == 1 ? y1 : 1.0);
                            v[uu] = fma(sTp[t2] * h1, h2, v[uu]);
                        }
                    }
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) sA[(4 * fw + uu) * GM_LDA + k] = v[uu];
                }
                if (ft == 0) { sHdr[b * 4] = kc4; sHdr[b * 4 + 1] = (ch == nchunks - 1); }
                kdone += kc;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sBar[b]);            // this warp's share of the buffer is published
                WS_T(4);
#ifndef WS_NO_PREFETCH
                if (ch == 0 && fw == 0 && tnext < tile_end) {
                    // the next tile's terms are streamed from DRAM once (irregular books): pull their lines into L2 a tile ahead,
                    // two lanes per unit
                    const int pu = lane >> 1;
                    const int64_t o = __shfl_sync(0xffffffffu, nx_off, pu);
                    const int puid = __shfl_sync(0xffffffffu, nx_uid, pu);
                    if (puid >= 0) {
                        const char* w0 = reinterpret_cast<const char*>(a.weight) + o * 16;
                        const char* a0 = reinterpret_cast<const char*>(a.term_p ? a.term_p : a.amt) + o * 8;
                        const char* n0 = reinterpret_cast<const char*>(a.node) + o * 8;
                        const int64_t skip = (lane & 1) * 128;
                        for (const char* q = w0 - (reinterpret_cast<uintptr_t>(w0) & 127) + skip; q < w0 + (int64_t)nP * 16; q += 256)
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
                        for (const char* q = a0 - (reinterpret_cast<uintptr_t>(a0) & 127) + skip; q < a0 + (int64_t)nP * 8; q += 256)
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
                        if (!a.term_p)
                        for (const char* q = n0 - (reinterpret_cast<uintptr_t>(n0) & 127) + skip; q < n0 + (int64_t)nP * 8; q += 256)
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
                    }
                }
#endif
                if (++bslot == GW_NBUF) { bslot = 0; ++buse; }
            }
        }
        WS_TFLUSH(0, ft == 0);
        return;
    }

    // ================================== mma warps: K loops and epilogues ==================================
    const int ng = tid >> 5;
    const int ar = lane >> 2, ac = lane & 3;
    const int oj = tid >> 3, ok4 = (tid & 7) * 4;        // epilogue: this thread owns gamma entries (oj, ok4..ok4+3)
    if (tid < 8) sStage[tid * SM::LDS_ + SM::NCS] = 0.0;
    double* my_tg = sTot + 33 + oj * CAV_RW + ok4;
    my_tg[0] = my_tg[1] = my_tg[2] = my_tg[3] = 0.0;
    if (tid <= CAV_RW) sTot[tid] = 0.0;

    int tslot = 0, bslot = 0, buse = 0;
    int n_out = 0;                                               // row segments this warp has sent (bulk-store slot = n_out % GW_NSTORE)
    WS_TDECL();
    for (int tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        WS_T0();
        const int tp = tslot;
        tslot = tslot + 1 == GW_NBUF ? 0 : tslot + 1;
        const int* sUnit = sUnitAll + tp * GT_TM;
        const int* sCol = sColAll + tp * SM::NCS;
        const int64_t* sOut = sOutAll + tp * GT_TM;
        const double* sW = sWall + tp * GT_TM;
        const double* sPv = sPvAll + tp * GT_TM;
        mbar_wait(&sBar[bslot], buse & 1);                       // first chunk of the tile (publishes the tile data as well)
        WS_T(0);
        const unsigned mask = sMask[tp];
        const int na = __popc(mask), nc = na * (na + 1) / 2, ncols = nc + na, nnt = (ncols + 7) >> 3;
        const int ntw = nnt > ng ? (nnt - ng + 7) >> 3 : 0;      // n-tiles of this warp: ng, ng+8, ...
        int cn[NT];
        double c[2][NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            cn[n] = (n < ntw) ? sCol[(ng + 8 * n) * 8 + ar] : (GT_NPACK + CAV_RW);
            c[0][n][0] = c[0][n][1] = c[1][n][0] = c[1][n][1] = 0.0;
        }
        bool last = false;
        while (!last) {
            const int b = bslot;
            WS_T(1);
            mbar_wait(&sBar[b], buse & 1);
            WS_T(0);
            const int kc4 = sHdr[b * 4];
            last = sHdr[b * 4 + 1] != 0;
            const double* sA = sAall + b * GT_TM * GM_LDA;
            const int* sRow = sRowAll + b * GM_KC;
#ifdef MMA_DIAG_NOMMA
            if (ntw > 0 && a.tile_units == nullptr) {
#else
            if (ntw > 0) {
#endif
                const double* A0 = sA + ar * GM_LDA + ac;
                const double* A1 = A0 + 8 * GM_LDA;
#pragma unroll kMmaUnroll
                for (int k = 0; k < kc4; k += 4) {
                    const double a0 = A0[k], a1 = A1[k];
                    const double* rowp = a.T + (size_t)sRow[k + ac] * GT_NC;
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        if (n < ntw) {                    // warp-uniform
#ifdef MMA_DIAG_BSMEM       // diagnostic build: B operands from shared memory, conflict-free pattern (wrong values; timing only)
                            const double bv = sAall[(k * 4 + ac * 4 + ar + n * 16) & 1023];
#else
                            const double bv = __ldg(rowp + cn[n]);
#endif
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(c[0][n][0]), "+d"(c[0][n][1]) : "d"(a0), "d"(bv));
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(c[1][n][0]), "+d"(c[1][n][1]) : "d"(a1), "d"(bv));
                        }
                    }
                }
            }
            if (!last) {                                         // hand the half back; the last chunk's goes back after the epilogue
                __syncwarp();
                if (lane == 0) mbar_arrive(&sBar[GW_NBUF + b]);
                if (++bslot == GW_NBUF) { bslot = 0; ++buse; }
            }
        }
        WS_T(1);
        // ---- epilogue: stage 8 units at a time, scatter to full symmetric rows ----
        int pk4[4], pd = SM::NCS;
        {   // real pillar r lives at position pos_of[r] of the permuted order the mask and the tables use
            const int qj = a.pp.pos_of[oj];
            const int ij = ((mask >> qj) & 1u) ? __popc(mask & ((1u << qj) - 1u)) : -1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int qk = a.pp.pos_of[ok4 + q];
                const int ik = ((mask >> qk) & 1u) ? __popc(mask & ((1u << qk) - 1u)) : -1;
                pk4[q] = (ij < 0 || ik < 0) ? SM::NCS : (ij >= ik ? ij * (ij + 1) / 2 + ik : ik * (ik + 1) / 2 + ij);
            }
            if (tid < CAV_RW) {
                const int qd = a.pp.pos_of[tid];
                if ((mask >> qd) & 1u) pd = nc + __popc(mask & ((1u << qd) - 1u));
            }
        }
        // portfolio partials of this tile: two interleaved accumulator sets (units alternate) - every DFMA here waits in the
        // FP64 pipe behind the other CTA's DMMAs, so the length of the dependent chain is what costs
        double tg[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}}, tdl[2] = {0.0, 0.0}, tpv[2] = {0.0, 0.0};
        WS_T(2);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            bar_named(1, 256);                           // stage rows free
            double* st = sStage + (size_t)ar * SM::LDS_ + 2 * ac;
#pragma unroll
            for (int n = 0; n < NT; ++n)
                if (n < ntw)
                    *reinterpret_cast<double2*>(st + (ng + 8 * n) * 8) = make_double2(c[mt][n][0], c[mt][n][1]);
            bar_named(1, 256);
#pragma unroll
            for (int s8 = 0; s8 < 8; ++s8) {
                const int u = mt * 8 + s8;
                const int uid = sUnit[u];
                if (uid >= 0) {
                    const double* row_s = sStage + (size_t)s8 * SM::LDS_;
                    const int64_t row = sOut[u];
                    const double W = sW[u];
                    const double g0 = row_s[pk4[0]], g1 = row_s[pk4[1]], g2 = row_s[pk4[2]], g3 = row_s[pk4[3]];
                    if (a.out_cgamma) {                   // compact unit row: the staged packed triangle as it is (coalesced)
                        double* cdst = a.out_cgamma + (size_t)uid * GT_NPACK;
                        for (int cc = tid; cc < nc; cc += 256) cdst[cc] = row_s[cc];
                        if (tid == 0) a.out_cmask[uid] = mask;
                    }
#ifdef MMA_DIAG_NOSTORE
                    if (a.out_gamma && g0 == 1.2345e-300) {
#else
                    if (a.out_gamma) {
#endif
#if GW_NSTORE > 0
                        // The warp's 4 matrix rows of this unit are 1 KB of contiguous output: staged in a warp-private
                        // buffer and sent with one bulk store (cp.async.bulk), which the copy engine drains while the warp
                        // goes on - plain stores stall the mma warps whenever the store path backs up (HBM write bursts of
                        // 128 KB per tile), measured as ~9k cycles of epilogue per tile.
                        double* rb = sRowBuf + (size_t)(ng * GW_NSTORE + n_out % GW_NSTORE) * 128;
                        if (n_out >= GW_NSTORE) {                // the slot's previous segment has been read out
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(GW_NSTORE - 1) : "memory");
                            __syncwarp();
                        }
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(smem_addr(rb + lane * 4)), "d"(g0), "d"(g1) : "memory");
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(smem_addr(rb + lane * 4 + 2)), "d"(g2), "d"(g3) : "memory");
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            double* dst = a.out_gamma + (size_t)row * CAV_RR + ng * 128;
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1024;"
                                         :: "l"(dst), "r"(smem_addr(rb)) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        ++n_out;
#else
                        double* dst = a.out_gamma + (size_t)row * CAV_RR + oj * CAV_RW + ok4;
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(dst), "d"(g0), "d"(g1), "d"(g2), "d"(g3) : "memory");
#endif
                    }
                    if (a.partials) {
                        tg[s8 & 1][0] = fma(W, g0, tg[s8 & 1][0]); tg[s8 & 1][1] = fma(W, g1, tg[s8 & 1][1]);
                        tg[s8 & 1][2] = fma(W, g2, tg[s8 & 1][2]); tg[s8 & 1][3] = fma(W, g3, tg[s8 & 1][3]);
                    }
                    if (tid < CAV_RW) {
                        const double dl = row_s[pd];
                        if (a.out_delta) a.out_delta[(size_t)row * CAV_RW + tid] = dl;
                        tdl[s8 & 1] = fma(W, dl, tdl[s8 & 1]);
                    }
                    if (tid == 32) {
                        if (a.out_pv) a.out_pv[row] = sPv[u];
                        tpv[s8 & 1] = fma(W, sPv[u], tpv[s8 & 1]);
                    }
                }
            }
        }
        if (a.partials) {
#pragma unroll
            for (int q = 0; q < 4; ++q) my_tg[q] += tg[0][q] + tg[1][q];
            if (tid < CAV_RW) sTot[1 + tid] += tdl[0] + tdl[1];
            if (tid == 32) sTot[0] += tpv[0] + tpv[1];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sBar[GW_NBUF + bslot]);      // the tile's last buffer, and with it the tile slot, go back
        WS_T(3);
        if (++bslot == GW_NBUF) { bslot = 0; ++buse; }
    }
#if GW_NSTORE > 0
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // shared memory stays put until the last segment is out
#endif
    WS_TFLUSH(8, tid == 0);
    if (a.partials) {     // this CTA's partial row (every launch owns its own block of rows and overwrites it)
        double* Pr = a.partials + (size_t)blockIdx.x * CAV_NOUT;
#pragma unroll
        for (int q = 0; q < 4; ++q) Pr[33 + oj * CAV_RW + ok4 + q] = my_tg[q];
        if (tid < CAV_RW) Pr[1 + tid] = sTot[1 + tid];
        if (tid == 32) Pr[0] = sTot[0];
    }
}

// ------------------------------------------------------------------------------------------
// XccyCurve bootstrap (cavour/trades/rates/xccy_curve.py:954-1206) with exact forward-mode tangents w.r.t. the pillar basis
// spreads: values, d(DF)/d(spread) (the reference's _jac_basis, jacrev through the scan, xccy_curve.py:594) and
// d2(DF)/d(spread)^2 (_hess_basis, xccy_curve.py:603-606), batched over shocked spread sets.
// The scan is serial in the payment points (each depends on the previous XCCY node and, at a maturity, on the swap's earlier
// cashflows), so the parallel axes are the derivative entries and the scenarios: one single-warp CTA per (Hessian row a,
// scenario), lane b = column.  A warp carries everything its row needs - the values, the gradient entry of lane b and of
// row a (a shuffle), the Hessian entry (a, b) - through the scan on its own: no barrier, the per-swap sums of known
// cashflow PVs (value, gradient, Hessian row) live in shared memory.  ORDER 0 / 1 run one warp per scenario.
//   d_int = DF_prev * grow,  grow = DF_ois(t)/DF_ois(t_prev) * exp(-s_k (t - t_prev))          (s_k: spread of the point's swap)
//   d grow = -(dt) grow e_k,  d2 grow = dt^2 grow e_k e_k^T;  cash = base + s_k * sens,  pv = cash * d_int
//   at a maturity DF = num / den, num = spot * known - pv_dom, den = -spot * cash:
//   dDF = (dnum - DF dden)/den,  d2DF = (d2num - dDF (x) dden - dden (x) dDF)/den.
// ------------------------------------------------------------------------------------------
#define XC_EXCH 1
#define XC_ATVAL 4
#define XC_MAT 8

template <int ORDER>
__global__ void __launch_bounds__(32)
k_xccy_scan(int n_pts, int nb, const double* __restrict__ pt_time, const int* __restrict__ pt_swap, const int* __restrict__ pt_flags,
            const double* __restrict__ pt_sens, const double* __restrict__ pt_base, const double* __restrict__ pt_dfois,
            const double* __restrict__ pt_pvdom, double spot, const double* __restrict__ spreads,
            double* __restrict__ df_out, double* __restrict__ jac_out, double* __restrict__ hess_out)
{
    extern __shared__ double xs[];
    double* kV = xs;                       // [nb]      known cashflow PVs per swap
    double* kG = kV + 32;                  // [nb][32]  their gradients (lane b)
    double* kH = kG + 32 * 32;             // [nb][32]  row a of their Hessians
    const int a = blockIdx.x, sc = blockIdx.y, b = threadIdx.x;
    const double* s = spreads + (size_t)sc * nb;
    for (int k = 0; k < nb; ++k) { if (b == 0) kV[k] = 0.0; kG[k * 32 + b] = 0.0; if (ORDER > 1) kH[k * 32 + b] = 0.0; }
    __syncwarp();
    double dfp = 1.0, gp = 0.0, hp = 0.0, t_prev = 0.0, o_prev = 1.0;
    bool have_prev = false;
    for (int i = 0; i < n_pts; ++i) {
        const int k = pt_swap[i], fl = pt_flags[i];
        const double t = pt_time[i], sens = pt_sens[i], o = pt_dfois[i], sk = s[k];
        const double cash = pt_base[i] + sk * sens;
        const double eb = (b == k) ? 1.0 : 0.0, ea = (a == k) ? 1.0 : 0.0;
        const double dcb = sens * eb, dca = sens * ea;
        const double dt = have_prev ? t - t_prev : t;
        const double grow = (have_prev ? o / o_prev : o) * exp(-sk * dt);
        const double d_int = dfp * grow;                                   // dfp = 1 before the first node
        const double gpa = ORDER > 1 ? __shfl_sync(0xffffffffu, gp, a) : 0.0;
        const double g_int = gp * grow - dt * d_int * eb;
        const double g_inta = gpa * grow - dt * d_int * ea;
        double h_int = 0.0;
        if (ORDER > 1) h_int = hp * grow - dt * grow * (gpa * eb + ea * gp) + dt * dt * d_int * ea * eb;
        double pvV = 0.0, pvG = 0.0, pvH = 0.0;
        if (fl & XC_ATVAL) { pvV = cash; pvG = dcb; }
        else if (!(fl & XC_MAT)) {
            pvV = cash * d_int;
            pvG = dcb * d_int + cash * g_int;
            if (ORDER > 1) pvH = dca * g_int + g_inta * dcb + cash * h_int;
        }
        double df = d_int, g = g_int, h = h_int;
        if (fl & XC_MAT) {
            const double known = kV[k] + pvV, knownG = kG[k * 32 + b] + pvG;
            const double num = -(pt_pvdom[i] + spot * (-known)), den = spot * (-cash);
            if (fabs(den) > 1e-12) {
                const double dnum = spot * knownG, dden = -spot * dcb;
                df = num / den;
                g = (dnum - df * dden) / den;
                if (ORDER > 1) {
                    const double ga = __shfl_sync(0xffffffffu, g, a), ddena = -spot * dca;
                    h = (spot * (kH[k * 32 + b] + pvH) - ga * dden - ddena * g) / den;
                }
            }
        } else {
            __syncwarp();
            if (b == 0) kV[k] += pvV;
            kG[k * 32 + b] += pvG;
            if (ORDER > 1) kH[k * 32 + b] += pvH;
            __syncwarp();
        }
        if (a == 0) {
            if (b == 0) df_out[(size_t)sc * n_pts + i] = df;
            if (ORDER > 0 && jac_out && b < nb) jac_out[((size_t)sc * n_pts + i) * nb + b] = g;
        }
        if (ORDER > 1 && hess_out && b < nb) hess_out[(((size_t)sc * n_pts + i) * nb + a) * nb + b] = h;
        if (!(fl & XC_ATVAL)) { dfp = df; gp = g; hp = h; t_prev = t; o_prev = o; have_prev = true; }
    }
}

// totals[e] = sum_rows partials[row][e]; one CTA per entry: threads stride the rows, a fixed butterfly per warp, the eight
// warp sums added in warp order (bitwise reproducible for a given grid).  (A warp per entry left 33 warps on the whole GPU
// for the PV + delta request, each waiting on ~300 strided loads: 43 us for 2.5 MB.)
__global__ void __launch_bounds__(256)
k_reduce_partials(const double* __restrict__ partials, int64_t n_rows, double* totals, int n_entries)
{
    __shared__ double s_w[8];
    const int e = blockIdx.x;
    if (e >= n_entries) { if (threadIdx.x == 0) totals[e] = 0.0; return; }      // entries the units stage did not produce (no gamma)
    double s = 0.0;
    for (int64_t w = threadIdx.x; w < n_rows; w += 256) s += partials[w * CAV_NOUT + e];
    const double t = block_sum_fixed(s, s_w);
    if (threadIdx.x == 0) totals[e] = t;
}

// ------------------------------------------------------------------------------------------
// Expansion: trade = sum_k weight_k * unit_k.  One CTA per group (trades sharing unit ids);
// each thread keeps its slice of the group's unit outputs in registers and streams one 8 KB
// gamma row per trade with 16-byte stores.  This is the HBM-write-bound stage.
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256)
k_expand(const int64_t* __restrict__ group_offsets, const int* __restrict__ group_units,
         const double* __restrict__ comp_weight, const int64_t* __restrict__ out_index,
         const double* __restrict__ u_pv, const double* __restrict__ u_delta, const double* __restrict__ u_gamma,
         double* pv, double* delta, double* gamma)
{
    __shared__ double s_w[256][K];
    __shared__ int64_t s_row[256];
    const int gidx = blockIdx.x, tid = threadIdx.x;
    const int64_t t0 = group_offsets[gidx];
    const int cnt = (int)(group_offsets[gidx + 1] - t0);        // <= 256 (checked at upload)
    if (tid < cnt) {
#pragma unroll
        for (int k = 0; k < K; ++k) s_w[tid][k] = comp_weight[(t0 + tid) * K + k];
        s_row[tid] = out_index ? out_index[t0 + tid] : t0 + tid;
    }
    int uid[K];
#pragma unroll
    for (int k = 0; k < K; ++k) uid[k] = group_units[gidx * K + k];
    double2 ga[K], gb[K];
    double dl[K], pvv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (gamma) {
            const double2* src = reinterpret_cast<const double2*>(u_gamma + (size_t)uid[k] * CAV_RR);
            ga[k] = src[tid * 2];
            gb[k] = src[tid * 2 + 1];
        }
        dl[k] = (delta && tid < 32) ? u_delta[(size_t)uid[k] * CAV_RW + tid] : 0.0;
        pvv[k] = (tid == 32) ? u_pv[uid[k]] : 0.0;
    }
    // output rows are written once and never read by this kernel: evict_first keeps them from pushing the unit rows the
    // next groups are about to read out of L2 (store microbenchmark: 6.45 -> 6.50 TB/s)
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
        double w[K];
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = s_w[i][k];
        const int64_t row = s_row[i];
        if (gamma) {
            double2 a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                a.x += w[k] * ga[k].x; a.y += w[k] * ga[k].y;
                b.x += w[k] * gb[k].x; b.y += w[k] * gb[k].y;
            }
            // one 32-byte store per thread: a warp instruction covers 1 KB of the row contiguously
            // (two 16-byte stores per thread leave 16-byte gaps per instruction and halve the rate)
            double* dst = gamma + (size_t)row * CAV_RR + tid * 4;
            asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1, %2, %3, %4}, %5;"
                         :: "l"(dst), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y), "l"(pol) : "memory");
        }
        if (delta && tid < 32) {
            double sdl = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) sdl += w[k] * dl[k];
            delta[(size_t)row * CAV_RW + tid] = sdl;
        }
        if (pv && tid == 32) {
            double spv = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) spv += w[k] * pvv[k];
            pv[row] = spv;
        }
    }
}

// Gamma expansion from COMPACT unit rows.  The tile kernels leave a unit's gamma as the packed triangle over the active
// pillars of its tile (u_cgamma[unit][<= 528], u_cmask[unit]; ~30 MB instead of 205 MB for the 25 100 units of the 1M-trade
// book), written right before this kernel runs, so the expansion's unit reads are L2 hits instead of DRAM reads threaded
// through the 8 GB write stream (store microbenchmark tools/expand_bench.cu, variants F / K: 6.50 -> 6.90 TB/s).  Thread t
// owns matrix entries (t >> 3, 4 (t & 7) ..+3) as in k_expand; their positions in the packed triangle follow from the mask and
// the pillar permutation exactly as in the tile kernel's epilogue, so every row is bit-identical to the full-row path.
template <int K>
__global__ void __launch_bounds__(256)
k_expand_c(const int64_t* __restrict__ group_offsets, const int* __restrict__ group_units,
           const double* __restrict__ comp_weight, const int64_t* __restrict__ out_index,
           const double* __restrict__ u_cgamma, const unsigned* __restrict__ u_cmask, PillarPerm pp, double* gamma)
{
    __shared__ double s_w[256][K];
    __shared__ int64_t s_row[256];
    const int gidx = blockIdx.x, tid = threadIdx.x;
    const int64_t t0 = group_offsets[gidx];
    const int cnt = (int)(group_offsets[gidx + 1] - t0);        // <= 256 (checked at upload)
    if (tid < cnt) {
#pragma unroll
        for (int k = 0; k < K; ++k) s_w[tid][k] = comp_weight[(t0 + tid) * K + k];
        s_row[tid] = out_index ? out_index[t0 + tid] : t0 + tid;
    }
    const int oj = tid >> 3, ok4 = (tid & 7) * 4;
    const int qj = pp.pos_of[oj];
    int qk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) qk[q] = pp.pos_of[ok4 + q];
    double gv[K][4];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int uid = group_units[gidx * K + k];
        const unsigned mask = __ldg(u_cmask + uid);
        const double* src = u_cgamma + (size_t)uid * GT_NPACK;
        const int ij = ((mask >> qj) & 1u) ? __popc(mask & ((1u << qj) - 1u)) : -1;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ik = ((mask >> qk[q]) & 1u) ? __popc(mask & ((1u << qk[q]) - 1u)) : -1;
            const int idx = ij >= ik ? ij * (ij + 1) / 2 + ik : ik * (ik + 1) / 2 + ij;
            gv[k][q] = (ij < 0 || ik < 0) ? 0.0 : __ldg(src + idx);
        }
    }
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
        double w[K];
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = s_w[i][k];
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) { x0 += w[k] * gv[k][0]; x1 += w[k] * gv[k][1]; x2 += w[k] * gv[k][2]; x3 += w[k] * gv[k][3]; }
        double* dst = gamma + (size_t)s_row[i] * CAV_RR + tid * 4;
        asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1, %2, %3, %4}, %5;"
                     :: "l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3), "l"(pol) : "memory");
    }
}

// PV and delta rows of the expansion, gathered per OUTPUT row (lane = pillar) from the unit results with the
// row-ordered tables of k_row_tables: consecutive rows write consecutive memory, instead of the 8-byte /
// 256-byte scatters the group-ordered k_expand would issue (measured: 1M scattered 8-byte PV stores alone cost
// 0.1 ms).  A warp owns 8 consecutive rows: it fetches their (unit, weight) pairs with one coalesced load,
// issues all unit-row gathers before the first store (the kernel is latency-, not bandwidth-bound otherwise)
// and writes the 8 PVs as one 64-byte run.
#define XR_ROWS 8
template <int K>
__global__ void __launch_bounds__(256)
k_expand_rows(int64_t n_trades, const int* __restrict__ row_units, const double* __restrict__ row_weight,
              const double* __restrict__ u_pv, const double* __restrict__ u_delta, double* pv, double* delta)
{
    const int lane = threadIdx.x & 31;
    const int64_t row0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * XR_ROWS;
    if (row0 >= n_trades) return;
    const int nr = (n_trades - row0) < XR_ROWS ? (int)(n_trades - row0) : XR_ROWS;
    int my_u = 0;
    double my_w = 0.0;
    if (lane < nr * K) { my_u = __ldg(row_units + row0 * K + lane); my_w = __ldg(row_weight + row0 * K + lane); }
    double d[XR_ROWS], p = 0.0;
#pragma unroll
    for (int r = 0; r < XR_ROWS; ++r) {
        d[r] = 0.0;
        double pr = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int u = __shfl_sync(0xffffffffu, my_u, r * K + k);
            const double w = __shfl_sync(0xffffffffu, my_w, r * K + k);
            if (delta) d[r] = fma(w, __ldg(u_delta + (size_t)u * CAV_RW + lane), d[r]);
            if (pv) pr = fma(w, __ldg(u_pv + u), pr);
        }
        if (lane == r) p = pr;
    }
    if (delta) {
#pragma unroll
        for (int r = 0; r < XR_ROWS; ++r)
            if (r < nr) delta[(row0 + r) * CAV_RW + lane] = d[r];
    }
    if (pv && lane < nr) pv[row0 + lane] = p;
}

// Same rows with 16-byte accesses: a half-warp owns a row (lane = two adjacent pillars), a warp 16 consecutive rows, two
// rows per step.  Half the instructions of k_expand_rows for the same bytes (that kernel is latency-, not bandwidth-bound:
// 264 MB in 86 us; this one 64 us); a warp store covers 512 contiguous bytes.
#define XR2_ROWS 16
template <int K>
__global__ void __launch_bounds__(256)
k_expand_rows2(int64_t n_trades, const int* __restrict__ row_units, const double* __restrict__ row_weight,
               const double* __restrict__ u_pv, const double* __restrict__ u_delta, double* pv, double* delta)
{
    const int lane = threadIdx.x & 31, half = lane >> 4, j = lane & 15;
    const int64_t row0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * XR2_ROWS;
    if (row0 >= n_trades) return;
    const int nr = (n_trades - row0) < XR2_ROWS ? (int)(n_trades - row0) : XR2_ROWS;
    // (unit, weight) of the warp's rows: entry e = r * K + k lives in lane e & 31 of slot e >> 5
    constexpr int NSLOT = (XR2_ROWS * K + 31) / 32;
    int my_u[NSLOT];
    double my_w[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int e = s * 32 + lane;
        const bool in = e < nr * K;
        my_u[s] = in ? __ldg(row_units + row0 * K + e) : 0;
        my_w[s] = in ? __ldg(row_weight + row0 * K + e) : 0.0;
    }
    double p_keep = 0.0;
#pragma unroll
    for (int it = 0; it < XR2_ROWS / 2; ++it) {
        const int r = 2 * it + half;
        double2 d = make_double2(0.0, 0.0);
        double pr = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int e = r * K + k;
            int u = 0;
            double w = 0.0;
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) {          // (per-lane source lane: the two half-warps read different entries)
                const int us = __shfl_sync(0xffffffffu, my_u[s], e & 31);
                const double ws = __shfl_sync(0xffffffffu, my_w[s], e & 31);
                if ((e >> 5) == s) { u = us; w = ws; }
            }
            if (delta) {
                const double2 g = __ldg(reinterpret_cast<const double2*>(u_delta + (size_t)u * CAV_RW) + j);
                d.x = fma(w, g.x, d.x);
                d.y = fma(w, g.y, d.y);
            }
            if (pv) pr = fma(w, __ldg(u_pv + u), pr);
        }
        if (delta && r < nr) *reinterpret_cast<double2*>(delta + (row0 + r) * CAV_RW + 2 * j) = d;
        if (j == it) p_keep = pr;                      // lane (half, j) keeps the PV of row 2 j + half
    }
    if (pv && j < XR2_ROWS / 2 && 2 * j + half < nr) pv[row0 + 2 * j + half] = p_keep;
}

// ------------------------------------------------------------------------------------------
// Chain rule as a batched FP64 tensor-core GEMM (the reference's `jnp.dot(grad_dfs, jac)`,
// engine.py:2554/2912, for all units at once):
//     delta[U][32] = Q[U][Gp] * g[Gp][32],   Q[u][n] = dPV_u / d ln d_n = sum_terms p * w
// k_node_grad builds the dense node-gradient rows, k_chain_gemm_dmma contracts them with the
// (1e-4-scaled) log-DF Jacobian using mma.sync.m8n8k4.f64 (DMMA; there is no FP64 tcgen05 kind).
// This is the dense formulation (2*Gp*32 flops per unit, 2.6x the sparse per-cashflow chain of
// k_units) and is kept as a measured alternative, not the default path.
// ------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(256)
k_node_grad(int64_t n_units, int Gp, const int64_t* __restrict__ unit_offsets, const double* __restrict__ amt,
            const double* __restrict__ weight, const int* __restrict__ node, const double* __restrict__ L,
            double* Q, double* unit_pv)
{
    const int lane = threadIdx.x & 31;
    const int64_t u = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (u >= n_units) return;
    double* q = Q + (size_t)u * Gp;
    for (int n = lane; n < Gp; n += 32) q[n] = 0.0;
    __syncwarp();
    double pv = 0.0;
    for (int64_t i = unit_offsets[u] + lane; i < unit_offsets[u + 1]; i += 32) {
        double ell = 0.0, w[NP];
        int nn[NP];
#pragma unroll
        for (int m = 0; m < NP; ++m) { w[m] = weight[i * NP + m]; nn[m] = node[i * NP + m]; ell += w[m] * L[nn[m]]; }
        const double p = amt[i] * exp(ell);
        pv += p;
#pragma unroll
        for (int m = 0; m < NP; ++m)
            if (w[m] != 0.0) atomicAdd(q + nn[m], p * w[m]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pv += __shfl_xor_sync(0xffffffffu, pv, o);
    if (lane == 0 && unit_pv) unit_pv[u] = pv;
}

#define CAV_GEMM_KC 256          // nodes staged per shared-memory chunk (multiple of 16)
#define CAV_GEMM_LD 33           // padded row stride of the staged g chunk (conflict-free B fragments)
// Each lane fetches its A operands with one 32-byte load per 16 nodes: lane (ar, ac) reads
// Q[unit ar][16t + 4ac .. 4ac+3], a full 128-byte line per unit row per warp instruction.  The
// K index of MMA step s is therefore the permuted node 16t + 4ac + s, and the B fragment is read
// from the same permuted row, so the product is unchanged.  Gp (row stride of Q) is a multiple of 16.
__global__ void __launch_bounds__(256)
k_chain_gemm_dmma(int64_t n_units, int Gp, const double* __restrict__ Q, const double* __restrict__ g, int G,
                  double* delta)
{
    extern __shared__ double s_g[];                     // [KC][LD]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t u0 = ((int64_t)blockIdx.x * 8 + wib) * 16;     // 16 units per warp (two m8 tiles)
    const int ar = lane >> 2, ac = lane & 3;             // fragment coordinates
    double c[2][4][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) c[m][n][0] = c[m][n][1] = 0.0;
    const int64_t ua = u0 + ar, ub = u0 + 8 + ar;
    const bool va = ua < n_units, vb = ub < n_units;
    const double* qa = Q + (size_t)(va ? ua : 0) * Gp + 4 * ac;
    const double* qb = Q + (size_t)(vb ? ub : 0) * Gp + 4 * ac;
    for (int k0 = 0; k0 < Gp; k0 += CAV_GEMM_KC) {
        const int kc = (Gp - k0) < CAV_GEMM_KC ? (Gp - k0) : CAV_GEMM_KC;
        __syncthreads();
        for (int e = threadIdx.x; e < kc * 32; e += 256) {
            const int r = e >> 5, col = e & 31;
            s_g[r * CAV_GEMM_LD + col] = (k0 + r < G) ? g[(size_t)(k0 + r) * 32 + col] : 0.0;
        }
        __syncthreads();
        if (u0 < n_units) {
#pragma unroll 2
            for (int k = 0; k < kc; k += 16) {
                double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
                if (va) asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                                     : "=d"(a0[0]), "=d"(a0[1]), "=d"(a0[2]), "=d"(a0[3]) : "l"(qa + k0 + k));
                if (vb) asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                                     : "=d"(a1[0]), "=d"(a1[1]), "=d"(a1[2]), "=d"(a1[3]) : "l"(qb + k0 + k));
#pragma unroll
                for (int st = 0; st < 4; ++st) {
                    const double* brow = s_g + (k + 4 * ac + st) * CAV_GEMM_LD + ar;
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const double b = brow[n * 8];
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(c[0][n][0]), "+d"(c[0][n][1]) : "d"(a0[st]), "d"(b));
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(c[1][n][0]), "+d"(c[1][n][1]) : "d"(a1[st]), "d"(b));
                    }
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int64_t u = u0 + m * 8 + ar;
        if (u < n_units) {
#pragma unroll
            for (int n = 0; n < 4; ++n)
                *reinterpret_cast<double2*>(delta + (size_t)u * 32 + n * 8 + 2 * ac) = make_double2(c[m][n][0], c[m][n][1]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// df_ad: fwd_k = -ln(d_{k+1}/d_k)/(x_{k+1}-x_k); f = interp(t, x[:-1], fwd);
//        i0 = searchsorted(x, t, 'right') - 1; DF = d[i0] exp(-f (t - x[i0]))
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int upper_bound(const double* x, int n, double t)
{   // first index with x[i] > t
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (x[mid] <= t) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void k_df_ad(const double* __restrict__ x, const double* __restrict__ d, int n,
                        const double* __restrict__ t, int64_t m, double* out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const double tt = t[q];
    const int nf = n - 1;                       // forward-rate knots x[0..nf-1]
    auto fwd = [&](int k) { return -log(d[k + 1] / d[k]) / (x[k + 1] - x[k]); };
    double f;
    if (tt < x[0]) f = fwd(0);
    else if (tt > x[nf - 1]) f = fwd(nf - 1);
    else {
        int i = upper_bound(x, nf, tt);
        i = i < 1 ? 1 : (i > nf - 1 ? nf - 1 : i);
        const double f0 = fwd(i - 1), f1 = fwd(i);
        const double dx = x[i] - x[i - 1];
        f = (fabs(dx) <= 4.930380657631324e-32) ? f0 : f0 + ((tt - x[i - 1]) / dx) * (f1 - f0);
    }
    int i0 = upper_bound(x, n, tt) - 1;
    if (i0 < 0) i0 += n;                        // numpy/jax negative index wraps
    out[q] = d[i0] * exp(-f * (tt - x[i0]));
}

// ------------------------------------------------------------------------------------------
// Path-A discount factors (Interpolator._uinterpolate, interpolator.py:69-170): exact hit on node 0; i = first
// node with x[i] >= t (n when t lies beyond the last node); LINEAR_ZERO_RATES interpolates the zero rates
// (the first segment and the extrapolation are flat in the zero rate), FLAT_FWD_RATES interpolates ln DF
// (extrapolation continues the last segment).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double node_df_path_a(int method, const double* __restrict__ x, const double* __restrict__ d,
                                                 int n, double t)
{
    if (t == x[0]) return d[0];
    // the reference scans linearly (`while times[i] < t and i < n - 1`); its own OIS curves carry duplicate node
    // times that differ in the last ulp, so a binary search is not equivalent
    int i = 0;
    while (x[i] < t && i < n - 1) ++i;
    if (t > x[i]) i = n;
    if (method == 4) {                            // LINEAR_ZERO_RATES
        double z1, z2;
        int lo, hi;
        if (i == 1) { z1 = z2 = -log(d[1]) / x[1]; lo = 0; hi = 1; }
        else if (i < n) { z1 = -log(d[i - 1]) / x[i - 1]; z2 = -log(d[i]) / x[i]; lo = i - 1; hi = i; }
        else { z1 = z2 = -log(d[n - 1]) / x[n - 1]; lo = n - 2; hi = n - 1; }
        const double z = ((x[hi] - t) * z1 + (t - x[lo]) * z2) / (x[hi] - x[lo]);
        return exp(-z * t);
    }
    const int lo = i < n ? i - 1 : n - 2, hi = i < n ? i : n - 1;      // FLAT_FWD_RATES
    const double y1 = -log(d[lo]), y2 = -log(d[hi]);
    const double y = ((x[hi] - t) * y1 + (t - x[lo]) * y2) / (x[hi] - x[lo]);
    return exp(-y);
}

__global__ void k_curve_df(int method, const double* __restrict__ x, const double* __restrict__ d, int n,
                           const double* __restrict__ t, int64_t m, double* out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < m) out[q] = node_df_path_a(method, x, d, n, t[q]);
}

// pv[i] = sum over the trade's cashflows of amt * DF(t) / DF(t_value); warp per trade, lanes stride the
// cashflows, fixed butterfly.  The curve nodes (<= 128 typical) are staged in shared memory.
#define CF_MAX_NODES 1024
__global__ void __launch_bounds__(256)
k_cashflow_pv(int method, const double* __restrict__ x, const double* __restrict__ d, int n, double t_value,
              int64_t n_trades, const int64_t* __restrict__ offsets, const double* __restrict__ t,
              const double* __restrict__ amt, double* pv)
{
    extern __shared__ double s_nodes[];           // x[n] | d[n]
    double* sx = s_nodes;
    double* sd = s_nodes + n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { sx[i] = x[i]; sd[i] = d[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t tr = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tr >= n_trades) return;
    const double inv0 = 1.0 / node_df_path_a(method, sx, sd, n, t_value);
    double acc = 0.0;
    for (int64_t c = offsets[tr] + lane; c < offsets[tr + 1]; c += 32) {
        const double tc = t[c];
        // (device-resident inputs are not scanned on the host: a negative time - the reference's LibError - gives NaN)
        acc += tc >= 0.0 ? (amt[c] * node_df_path_a(method, sx, sd, n, tc)) * inv0 : __longlong_as_double(0x7FF8000000000000ll);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) pv[tr] = acc;
}

// total = sum pv[i]: one CTA, fixed strides and butterfly (bitwise reproducible)
__global__ void __launch_bounds__(1024)
k_sum_fixed(const double* __restrict__ v, int64_t n, double* out)
{
    __shared__ double s[32];
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) a += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        a = s[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (threadIdx.x == 0) *out = a;
    }
}

// ------------------------------------------------------------------------------------------
// Scenarios: thread per scenario re-bootstraps the grid (DFs only) and stores ln DF;
// then unit PVs per scenario, then trades.
// ------------------------------------------------------------------------------------------
__global__ void k_scen_bootstrap(int G, int R, int n_scen, const double* __restrict__ rates /*[S][R]*/,
                                 const double* __restrict__ acc, const int* __restrict__ swap,
                                 const int* __restrict__ prev, double* Pbuf /*[G][S]*/, double* Ls /*[G][S]*/)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scen) return;
    for (int i = 0; i < G; ++i) {
        const int p = prev[i];
        const double r = rates[(size_t)s * R + swap[i]];
        const double a = acc[i];
        const double Pp = (p < 0) ? 0.0 : Pbuf[(size_t)p * n_scen + s];
        const double d = (1.0 - r * Pp) / (1.0 + r * a);
        Pbuf[(size_t)i * n_scen + s] = Pp + a * d;
        Ls[(size_t)i * n_scen + s] = log(d);
    }
}

// unit ids and weights per OUTPUT ROW (original trade order) from the group table, so that the row-ordered
// expansions write coalesced rows.  Warp per group, lanes stride its trades (a group has at most 256).
__global__ void __launch_bounds__(256)
k_row_tables(int64_t n_groups, int K, const int64_t* __restrict__ group_offsets,
             const int* __restrict__ group_units, const double* __restrict__ comp_weight,
             const int64_t* __restrict__ out_index, int* row_units, double* row_weight)
{
    const int64_t gi = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gi >= n_groups) return;
    const int64_t t0 = group_offsets[gi], t1 = group_offsets[gi + 1];
    for (int64_t t = t0 + (threadIdx.x & 31); t < t1; t += 32) {
        const int64_t row = out_index ? out_index[t] : t;
        for (int k = 0; k < K; ++k) {
            row_units[row * K + k] = group_units[gi * K + k];
            row_weight[row * K + k] = comp_weight[t * K + k];
        }
    }
}

// unit_pv[u][s]: block = 128 scenarios (threads) x loop over the unit's terms; grid (units, scen tiles)
template <int NP>
__global__ void __launch_bounds__(128)
k_scen_units(int n_scen, const int64_t* __restrict__ unit_offsets, const double* __restrict__ amt,
             const double* __restrict__ weight, const int* __restrict__ node,
             const double* __restrict__ Ls, double* unit_pv /*[U][S]*/)
{
    const int64_t u = blockIdx.x;
    const int s = blockIdx.y * blockDim.x + threadIdx.x;
    if (s >= n_scen) return;
    const int64_t t0 = unit_offsets[u], t1 = unit_offsets[u + 1];
    double pv = 0.0;
    for (int64_t i = t0; i < t1; ++i) {
        double ell = 0.0;
#pragma unroll
        for (int m = 0; m < NP; ++m) {
            const double w = weight[i * NP + m];
            if (w != 0.0) ell += w * Ls[(size_t)node[i * NP + m] * n_scen + s];
        }
        pv += amt[i] * exp(ell);
    }
    unit_pv[u * n_scen + s] = pv;
}

// Books with shared dates ask for the same discount factor many times per scenario (the 345k terms of the 100k-trade
// dedup book are 12.6k distinct (bracket, weights) queries).  DF cache path: every distinct query is evaluated once
// per scenario (k_scen_df, one exp each), the units then only gather and sum (k_scen_units_q).  dfq is [J][S] with the
// scenario index fastest; the grid runs query / unit fastest, so the 128-scenario slab of dfq a wave of CTAs reads
// (J x 1 KB) stays in L2.
__global__ void __launch_bounds__(128)
k_scen_df(int n_scen, const int2* __restrict__ q_node, const double2* __restrict__ q_w, const double* __restrict__ Ls,
          double* dfq)
{
    const int64_t j = blockIdx.x;
    const int s = blockIdx.y * blockDim.x + threadIdx.x;
    if (s >= n_scen) return;
    const int2 n = q_node[j];
    const double2 w = q_w[j];
    double ell = w.x * Ls[(size_t)n.x * n_scen + s];
    if (w.y != 0.0) ell += w.y * Ls[(size_t)n.y * n_scen + s];
    dfq[j * n_scen + s] = exp(ell);
}

__global__ void __launch_bounds__(128)
k_scen_units_q(int n_scen, const int64_t* __restrict__ unit_offsets, const double* __restrict__ amt,
               const int* __restrict__ term_q, const double* __restrict__ dfq, double* unit_pv /*[U][S]*/)
{
    const int64_t u = blockIdx.x;
    const int s = blockIdx.y * blockDim.x + threadIdx.x;
    if (s >= n_scen) return;
    const int64_t t0 = unit_offsets[u], t1 = unit_offsets[u + 1];
    double pv = 0.0;
    int64_t i = t0;
    for (; i + 4 <= t1; i += 4) {                 // four gathers in flight; summed in term order
        const double d0 = dfq[(size_t)term_q[i] * n_scen + s], d1 = dfq[(size_t)term_q[i + 1] * n_scen + s];
        const double d2 = dfq[(size_t)term_q[i + 2] * n_scen + s], d3 = dfq[(size_t)term_q[i + 3] * n_scen + s];
        pv += amt[i] * d0; pv += amt[i + 1] * d1; pv += amt[i + 2] * d2; pv += amt[i + 3] * d3;
    }
    for (; i < t1; ++i) pv += amt[i] * dfq[(size_t)term_q[i] * n_scen + s];
    unit_pv[u * n_scen + s] = pv;
}

// Even scenario counts: a thread owns two adjacent scenarios and gathers them with one 16-byte load, so the uniform
// loads (term_q, amt), the address arithmetic and the loop control are shared by two sums (the scalar kernel is issue
// bound: ncu issue active 70 %).  Same additions in the same order per scenario: bit-identical results.
__global__ void __launch_bounds__(128)
k_scen_units_q2(int n_scen, const int64_t* __restrict__ unit_offsets, const double* __restrict__ amt,
                const int* __restrict__ term_q, const double* __restrict__ dfq, double* unit_pv /*[U][S]*/)
{
    const int64_t u = blockIdx.x;
    const int s = 2 * (blockIdx.y * blockDim.x + threadIdx.x);
    if (s >= n_scen) return;
    const int64_t t0 = unit_offsets[u], t1 = unit_offsets[u + 1];
    const double* base = dfq + s;
    double2 pv = make_double2(0.0, 0.0);
    int64_t i = t0;
    for (; i + 4 <= t1; i += 4) {                 // four gathers in flight; summed in term order
        const double2 d0 = *reinterpret_cast<const double2*>(base + (size_t)term_q[i] * n_scen);
        const double2 d1 = *reinterpret_cast<const double2*>(base + (size_t)term_q[i + 1] * n_scen);
        const double2 d2 = *reinterpret_cast<const double2*>(base + (size_t)term_q[i + 2] * n_scen);
        const double2 d3 = *reinterpret_cast<const double2*>(base + (size_t)term_q[i + 3] * n_scen);
        const double a0 = amt[i], a1 = amt[i + 1], a2 = amt[i + 2], a3 = amt[i + 3];
        pv.x += a0 * d0.x; pv.y += a0 * d0.y; pv.x += a1 * d1.x; pv.y += a1 * d1.y;
        pv.x += a2 * d2.x; pv.y += a2 * d2.y; pv.x += a3 * d3.x; pv.y += a3 * d3.y;
    }
    for (; i < t1; ++i) {
        const double2 d = *reinterpret_cast<const double2*>(base + (size_t)term_q[i] * n_scen);
        const double a = amt[i];
        pv.x += a * d.x; pv.y += a * d.y;
    }
    *reinterpret_cast<double2*>(unit_pv + u * n_scen + s) = pv;
}

// pnl[s][row] = sum_k w_row,k unit_pv[u_row,k][s]; a 64x64 (rows x scenarios) tile is read with the scenario
// index fastest (256-byte runs of unit_pv, 16 independent gathers per thread) and written transposed with the row
// index fastest (512-byte runs of a P&L row).  The smaller 32x32 tile spent its time in barriers and exposed
// load latency (ncu: long scoreboard 17.9, barrier 5.2 warps per issue at 33 % of the DRAM peak).
#define SX_T 64
template <int K>
__global__ void __launch_bounds__(256)
k_scen_expand(int n_scen, int64_t n_trades, const int* __restrict__ row_units /*[N][K]*/,
              const double* __restrict__ row_weight, const double* __restrict__ unit_pv, double* pnl)
{
    extern __shared__ double tile[];                         // [SX_T][SX_T + 1]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int64_t rbase = (int64_t)blockIdx.x * SX_T;
    const int sbase = blockIdx.y * SX_T;
#pragma unroll
    for (int rr = 0; rr < SX_T / 8; ++rr) {
        const int r = ty + 8 * rr;
        const int64_t row = rbase + r;
        double v0 = 0.0, v1 = 0.0;
        if (row < n_trades) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double w = __ldg(row_weight + row * K + k);
                if (w != 0.0) {
                    const double* up = unit_pv + (size_t)__ldg(row_units + row * K + k) * n_scen + sbase + tx;
                    if (sbase + tx < n_scen) v0 = fma(w, up[0], v0);
                    if (sbase + 32 + tx < n_scen) v1 = fma(w, up[32], v1);
                }
            }
        }
        tile[r * (SX_T + 1) + tx] = v0;
        tile[r * (SX_T + 1) + 32 + tx] = v1;
    }
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < SX_T / 8; ++cc) {
        const int c = ty + 8 * cc;
        const int sc = sbase + c;
        if (sc >= n_scen) continue;
        double* dst = pnl + (size_t)sc * n_trades + rbase;
        if (rbase + tx < n_trades) __stcs(dst + tx, tile[tx * (SX_T + 1) + c]);
        if (rbase + 32 + tx < n_trades) __stcs(dst + 32 + tx, tile[(32 + tx) * (SX_T + 1) + c]);
    }
}

// Variant for an even number of scenarios: a 32 x 128 (rows x scenarios) tile, unit_pv read with 16-byte loads (a lane owns
// two adjacent scenarios in each half of the tile), so a row costs half the load instructions and its index / weight
// loads and address arithmetic are shared by four outputs instead of two (the 64x64 kernel issues ~67 warp instructions
// per 32 outputs and is half issue-bound, half latency-bound; ncu: issue active 52 %, long scoreboard 17.9).
#define SX2_R 32
#define SX2_S 128
template <int K>
__global__ void __launch_bounds__(256)
k_scen_expand2(int n_scen, int64_t n_trades, const int* __restrict__ row_units /*[N][K]*/,
               const double* __restrict__ row_weight, const double* __restrict__ unit_pv, double* pnl)
{
    __shared__ double tile[SX2_R * (SX2_S + 1)];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int64_t rbase = (int64_t)blockIdx.x * SX2_R;
    const int sbase = blockIdx.y * SX2_S;
    const int s0 = sbase + 2 * tx, s1 = s0 + 64;              // n_scen is even: a pair is inside or outside as a whole
    const bool in0 = s0 < n_scen, in1 = s1 < n_scen;
#pragma unroll
    for (int rr = 0; rr < SX2_R / 8; ++rr) {
        const int r = ty + 8 * rr;
        const int64_t row = rbase + r;
        double2 v0 = make_double2(0.0, 0.0), v1 = v0;
        if (row < n_trades) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double w = __ldg(row_weight + row * K + k);
                if (w != 0.0) {
                    const double* up = unit_pv + (size_t)__ldg(row_units + row * K + k) * n_scen;
                    if (in0) { const double2 a = *reinterpret_cast<const double2*>(up + s0); v0.x = fma(w, a.x, v0.x); v0.y = fma(w, a.y, v0.y); }
                    if (in1) { const double2 a = *reinterpret_cast<const double2*>(up + s1); v1.x = fma(w, a.x, v1.x); v1.y = fma(w, a.y, v1.y); }
                }
            }
        }
        double* t = tile + r * (SX2_S + 1) + 2 * tx;
        t[0] = v0.x; t[1] = v0.y; t[64] = v1.x; t[65] = v1.y;
    }
    __syncthreads();
    const bool row_in = rbase + tx < n_trades;
#pragma unroll 4
    for (int cc = 0; cc < SX2_S / 8; ++cc) {
        const int c = ty + 8 * cc;
        const int sc = sbase + c;
        if (sc < n_scen && row_in) __stcs(pnl + (size_t)sc * n_trades + rbase + tx, tile[tx * (SX2_S + 1) + c]);
    }
}

// ------------------------------------------------------------------------------------------
// Prefix chains of units.  Schedule units of one start date and growing maturity are prefixes of each other: the annuity
// of the 7Y swap is the annuity of the 6Y swap plus one term, same dates, same accruals.  The flattener numbers such units
// consecutively (classes are sorted by (effective, span)), so a chain is a run of units u0, u0+1, ... in which every term
// list starts with the whole list of its predecessor (k_scen_chain_flags: same DF query and bit-equal amount, position by
// position).  k_scen_units_chain then walks ONLY the last member's list and drops every member's value at its prefix
// point: the sum is the same left-to-right chain of FMAs k_scen_units_q2 runs per unit, so the unit values are bit-identical,
// at the gather work of one unit per chain instead of all of them (BASELINE config 4: 12.6k distinct queries against 339k
// terms - 27x fewer gathers from the L2-resident DF slab).  Units outside any run are chains of one.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_scen_chain_flags(int64_t n_units, const int64_t* __restrict__ unit_offsets, const double* __restrict__ amt,
                   const int* __restrict__ term_q, int* __restrict__ ext)
{
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    int e = 0;
    if (u > 0) {
        const int64_t p0 = unit_offsets[u - 1], t0 = unit_offsets[u], t1 = unit_offsets[u + 1];
        const int64_t lp = t0 - p0, lu = t1 - t0;
        if (lp > 0 && lu >= lp) {
            e = 1;
            for (int64_t i = 0; i < lp; ++i)
                if (term_q[p0 + i] != term_q[t0 + i] || __double_as_longlong(amt[p0 + i]) != __double_as_longlong(amt[t0 + i])) { e = 0; break; }
        }
    }
    ext[u] = e;
}

// DF query of a unit's last term (the key the chains are ordered by, see ensure_scen_chains)
__global__ void __launch_bounds__(256)
k_scen_unit_lastq(int64_t n_units, const int64_t* __restrict__ unit_offsets, const int* __restrict__ term_q, int* __restrict__ lastq)
{
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const int64_t t0 = unit_offsets[u], t1 = unit_offsets[u + 1];
    lastq[u] = t1 > t0 ? term_q[t1 - 1] : -1;
}

#define SCH_PER_CTA 8            // chains per CTA: half of a book's chains are two-term floating units, a CTA for each is all launch overhead
// chain_desc[c] = (first unit, members, terms of the last member, 0), chain_t0[c] = term offset of the last member: one load
// level instead of three (head -> unit offsets -> terms) in front of every chain's gathers, and the next chain's descriptor
// is fetched while this one is summed - the two-term floating units are pure latency otherwise.
__global__ void __launch_bounds__(128)
k_scen_units_chain(int n_scen, int n_chains, const int4* __restrict__ chain_desc, const int64_t* __restrict__ chain_t0,
                   const int64_t* __restrict__ unit_offsets, const double* __restrict__ amt, const int* __restrict__ term_q,
                   const double* __restrict__ dfq, double* unit_pv /*[U][S]*/)
{
    const int s = 2 * (blockIdx.y * blockDim.x + threadIdx.x);
    if (s >= n_scen) return;
    const double* base = dfq + s;
    const int c0 = blockIdx.x * SCH_PER_CTA;
    const int c_end = min(n_chains, c0 + SCH_PER_CTA);
    int4 nd = __ldg(chain_desc + c0);
    int64_t nt0 = __ldg(chain_t0 + c0);
    for (int c = c0; c < c_end; ++c) {
        const int u0 = nd.x, cnt = nd.y, L = nd.z;                    // the last member's list contains every member's
        const int64_t t0 = nt0;
        if (c + 1 < c_end) { nd = __ldg(chain_desc + c + 1); nt0 = __ldg(chain_t0 + c + 1); }
        double2 d[4];
        double a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                                 // four gathers in flight, refilled as they are consumed
            a[j] = 0.0; d[j] = make_double2(0.0, 0.0);
            if (j < L) { a[j] = __ldg(amt + t0 + j); d[j] = *reinterpret_cast<const double2*>(base + (size_t)__ldg(term_q + t0 + j) * n_scen); }
        }
        double2 pv = make_double2(0.0, 0.0);
        int m = 0;                                                    // next member to complete, at prefix length next_len
        int next_len = cnt == 1 ? L : (int)(unit_offsets[u0 + 1] - unit_offsets[u0]);
        while (m < cnt && next_len == 0) {                            // (no flattened book has empty units)
            *reinterpret_cast<double2*>(unit_pv + (size_t)(u0 + m) * n_scen + s) = pv;
            ++m;
            next_len = m < cnt ? (int)(unit_offsets[u0 + m + 1] - unit_offsets[u0 + m]) : -1;
        }
        for (int i = 0; i < L; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ii = i + j;
                if (ii < L) {
                    pv.x += a[j] * d[j].x; pv.y += a[j] * d[j].y;      // summed in term order, as k_scen_units_q2 does
                    if (ii + 4 < L) { a[j] = __ldg(amt + t0 + ii + 4); d[j] = *reinterpret_cast<const double2*>(base + (size_t)__ldg(term_q + t0 + ii + 4) * n_scen); }
                    while (ii + 1 == next_len) {
                        *reinterpret_cast<double2*>(unit_pv + (size_t)(u0 + m) * n_scen + s) = pv;
                        ++m;
                        next_len = m < cnt ? (int)(unit_offsets[u0 + m + 1] - unit_offsets[u0 + m]) : -1;
                    }
                }
            }
        }
    }
}

// Variant with bulk stores (even scenario and trade counts; the default): a 128 x 32 (trades x scenarios) tile assembled in
// shared memory in OUTPUT orientation (one row per scenario, trades contiguous) and written by the copy engine, one
// cp.async.bulk.global.shared::cta of up to 1 KB per scenario row.  A half-warp owns a trade pair: a thread gathers a 2 x 2 block
// (trades r, r+1 x scenarios s, s+1) with 16-byte reads of the unit values and leaves it with two 16-byte shared-memory
// stores; scenario row s starts at s*128 + 2*(s>>1) doubles, which keeps the rows 16-byte aligned and the quarter-warps of those
// stores on distinct banks.  Weights and unit ids are per trade, i.e. warp-uniform: lane j fetches those of the warp's j-th trade
// pair once per tile and the trips broadcast them with shuffles (one dependent load level instead of two per trip); a unit-row
// address is one 32 x 32 -> 64-bit multiply-add.  ~14 instructions per output in SASS against ~29 of k_scen_expand2.
// What the measurements say about this kernel family (profiles/r02m_experiments.txt): the P&L matrix can be written in exactly
// these tiles at the memset rate (7.5 TB/s, tools/pnl_store_bench.cu), so the store side is free; the gather side is bound by
// exposed load latency and by the SM's L2 -> L1 fill (two gathered bytes per byte written, trades arrive in trade order).  Hence
// the small tile: 33 KB of shared memory, four CTAs per SM, ~120 KB left to L1, where the unit rows of the schedules many trades
// share (spot starts: half of a typical book) survive across the CTAs of an SM.  5 or 6 CTAs per SM (less L1) are slower, a
// 128 x 64 tile (three CTAs, 58 KB of L1) is as slow as k_scen_expand2, a persistent double-buffered form (gather of tile i + 1
// beside the copy engine draining tile i) is slower still: resident warps and L1 beat overlap inside a CTA.
// Same sums in the same order as the other expansion kernels: bit-identical.
#define SX4_R 128
#define SX4_S 32
#define SX4_DOUBLES (SX4_S * SX4_R + SX4_S)
template <int K>
__global__ void __launch_bounds__(256, 4)
k_scen_expand4(int n_scen, int64_t n_trades, const int* __restrict__ row_units /*[N][K]*/,
               const double* __restrict__ row_weight, const double* __restrict__ unit_pv, double* pnl)
{
    extern __shared__ __align__(16) double tile4[];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int sp = tx & 15, half = tx >> 4;                   // scenario pair, trade pair of the warp's two
    const int64_t rbase = (int64_t)blockIdx.x * SX4_R;
    const int sbase = blockIdx.y * SX4_S;
    const int s0 = sbase + 2 * sp;
    const bool in0 = s0 < n_scen;                             // n_scen is even
    const int64_t left = n_trades - rbase;
    const int nloc = left < SX4_R ? (int)left : SX4_R;        // even, because n_trades is
    double* trow = tile4 + (2 * sp) * SX4_R + 2 * sp;          // row 2sp starts at 2sp*128 + 2sp doubles; row 2sp + 1 SX4_R further
    const char* ubase = reinterpret_cast<const char*>(unit_pv + (in0 ? s0 : 0));
    const unsigned row_bytes = (unsigned)n_scen * 8u;
    const double* wp = row_weight + rbase * K;
    const int* up = row_units + rbase * K;
    // trade pair of (trip it, half h) = 2 * (2 * (ty + 8 * it) + h); lane j < 8 prefetches the header of (it = j >> 1, h = j & 1)
    double2 hw0 = make_double2(0.0, 0.0), hw1 = hw0;
    int4 hu = make_int4(0, 0, 0, 0);
    if constexpr (K == 2) {
        const int j = tx & 7;
        const int rj = 2 * (2 * (ty + 8 * (j >> 1)) + (j & 1));
        if (rj < nloc) {
            hw0 = __ldg(reinterpret_cast<const double2*>(wp + 2 * rj));
            hw1 = __ldg(reinterpret_cast<const double2*>(wp + 2 * rj + 2));
            hu = __ldg(reinterpret_cast<const int4*>(up + 2 * rj));
        }
    }
#pragma unroll
    for (int it = 0; it < SX4_R / 32; ++it) {
        const int r = 2 * (2 * (ty + 8 * it) + half);
        double2 va = make_double2(0.0, 0.0), vb = va;
        double w[2 * K];
        int u[2 * K];
        if constexpr (K == 2) {
            const int src = 2 * it + half;
            w[0] = __shfl_sync(0xffffffffu, hw0.x, src); w[1] = __shfl_sync(0xffffffffu, hw0.y, src);
            w[2] = __shfl_sync(0xffffffffu, hw1.x, src); w[3] = __shfl_sync(0xffffffffu, hw1.y, src);
            u[0] = __shfl_sync(0xffffffffu, hu.x, src); u[1] = __shfl_sync(0xffffffffu, hu.y, src);
            u[2] = __shfl_sync(0xffffffffu, hu.z, src); u[3] = __shfl_sync(0xffffffffu, hu.w, src);
        }
        if (in0 && r < nloc) {
            if constexpr (K != 2) {
#pragma unroll
                for (int k = 0; k < 2 * K; ++k) { w[k] = __ldg(wp + r * K + k); u[k] = __ldg(up + r * K + k); }
            }
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (w[k] != 0.0) {
                    const double2 a = __ldg(reinterpret_cast<const double2*>(ubase + (unsigned long long)(unsigned)u[k] * row_bytes));
                    va.x = fma(w[k], a.x, va.x); va.y = fma(w[k], a.y, va.y);
                }
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (w[K + k] != 0.0) {
                    const double2 a = __ldg(reinterpret_cast<const double2*>(ubase + (unsigned long long)(unsigned)u[K + k] * row_bytes));
                    vb.x = fma(w[K + k], a.x, vb.x); vb.y = fma(w[K + k], a.y, vb.y);
                }
        }
        *reinterpret_cast<double2*>(trow + r) = make_double2(va.x, vb.x);
        *reinterpret_cast<double2*>(trow + SX4_R + r) = make_double2(va.y, vb.y);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < SX4_S) {
        const int c = threadIdx.x, sc = sbase + c;
        if (sc < n_scen) {
            unsigned long long pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            const double* src = tile4 + c * SX4_R + 2 * (c >> 1);
            double* dst = pnl + (size_t)sc * n_trades + rbase;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                         :: "l"(dst), "r"((unsigned)__cvta_generic_to_shared(src)), "r"((unsigned)nloc * 8u), "l"(pol) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

