/*
 * cavour_oracle.c - C restatement of the reference's per-trade OIS valuation + Greeks.
 *
 * TEST / BASELINE INFRASTRUCTURE - NOT PRODUCT CODE.  Built into oracle/liboracle.so and
 * used only by tests/ (cross-checked against oracle/cavour_oracle.py, which is pinned to
 * the reference's golden vectors) and by bench.py's cpu_baseline / --impl reference legs.
 *
 * It follows the reference's algorithm trade by trade, with no reuse across trades
 * (citations into /root/reference/cavour):
 *   interp()        InterpolatorAd.simple_interpolate         market/curves/interpolator_ad.py:210-243
 *                   (argmin over ALL nodes for the 1e-10 snap, +1e-12 shift, searchsorted right)
 *   leg terms       Engine._price_fixed_leg_jax                market/position/engine.py:2425-2448
 *                   Engine._float_leg_jax                      market/position/engine.py:2662-2728
 *   chain()         grad / dense G x G hessian w.r.t. node DFs, `g @ J * 1e-4`,
 *                   `(J.T H J + sum_k g_k C_k) * 1e-8`          engine.py:2551-2568, 2909-2926
 *   per trade       fixed analytics + floating analytics       engine.py:153-189
 * The curve tables (dfs, jac, hess) are passed in, i.e. computed ONCE per curve, which is
 * already more generous than the reference (it rebuilds them per Position, position.py:55).
 * dense=1 does the matrix products densely like the reference; dense=0 skips rows/columns
 * of the node-DF gradient/Hessian that are structurally zero (same arithmetic, fewer
 * multiplications by zero).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LZR 4
#define FF 1

typedef struct { int n[2]; double w[2]; int cnt; } bracket_t;

/* ln DF(t) = sum_k w_k ln d[n_k] */
static void interp(double t, const double* x, int G, int method, bracket_t* b)
{
    int best = 0; double bd = fabs(t - x[0]);
    for (int i = 1; i < G; ++i) { double di = fabs(t - x[i]); if (di < bd) { bd = di; best = i; } }
    if (bd < 1e-10) { b->cnt = 1; b->n[0] = best; b->w[0] = 1.0; return; }
    double ta = t + 1e-12;
    int lo = 0, hi = G;                    /* searchsorted(x, ta, side='right') */
    while (lo < hi) { int mid = (lo + hi) / 2; if (x[mid] <= ta) lo = mid + 1; else hi = mid; }
    int i = lo < 1 ? 1 : (lo > G - 1 ? G - 1 : lo);
    int a = i - 1, c = i;
    double c0, c1; int n0, n1, cnt;
    if (ta > x[G - 1]) { cnt = 1; n0 = G - 1; c0 = 1.0; n1 = 0; c1 = 0.0; }
    else if (ta < x[0]) { cnt = 1; n0 = 0; c0 = 1.0; n1 = 0; c1 = 0.0; }
    else {
        double dx = x[c] - x[a];
        if (fabs(dx) <= 4.930380657631324e-32) { cnt = 1; n0 = a; c0 = 1.0; n1 = 0; c1 = 0.0; }
        else { double w = (ta - x[a]) / dx; cnt = 2; n0 = a; c0 = 1.0 - w; n1 = c; c1 = w; }
    }
    if (method == LZR) {
        c0 = c0 * t / fmax(x[n0], 1e-15);
        if (cnt == 2) c1 = c1 * t / fmax(x[n1], 1e-15);
    }
    b->cnt = cnt; b->n[0] = n0; b->w[0] = c0; b->n[1] = n1; b->w[1] = c1;
}

/* add term c * prod DF^e to value / grad[G] / hess[G*G]; up to 3 brackets with exponents */
static double add_term(double c, const bracket_t* br, const double* ex, int nb, const double* d, int G,
                       double* grad, double* hess)
{
    int nodes[6]; double W[6]; int m = 0;
    double ell = 0.0;
    for (int q = 0; q < nb; ++q)
        for (int k = 0; k < br[q].cnt; ++k) {
            int n = br[q].n[k]; double w = br[q].w[k] * ex[q];
            ell += w * log(d[n]);
            int f = -1;
            for (int z = 0; z < m; ++z) if (nodes[z] == n) f = z;
            if (f < 0) { nodes[m] = n; W[m] = w; ++m; } else W[f] += w;
        }
    double p = c * exp(ell);
    if (grad) {
        for (int a = 0; a < m; ++a) {
            grad[nodes[a]] += p * W[a] / d[nodes[a]];
            if (hess) {
                for (int b = 0; b < m; ++b)
                    hess[(size_t)nodes[a] * G + nodes[b]] += p * W[a] * W[b] / (d[nodes[a]] * d[nodes[b]]);
                hess[(size_t)nodes[a] * G + nodes[a]] -= p * W[a] / (d[nodes[a]] * d[nodes[a]]);
            }
        }
    }
    return p;
}

/* delta += 1e-4 grad@J ; gamma += 1e-8 (J^T H J + sum_k grad_k C_k) */
static void chain(const double* grad, const double* hess, const double* J, const double* C, int G, int R, int dense,
                  double* delta, double* gamma, double* tmp /* G*R */)
{
    if (delta)
        for (int k = 0; k < G; ++k) {
            if (!dense && grad[k] == 0.0) continue;
            for (int r = 0; r < R; ++r) delta[r] += 1e-4 * grad[k] * J[(size_t)k * R + r];
        }
    if (!gamma) return;
    /* tmp = H J */
    memset(tmp, 0, sizeof(double) * G * R);
    for (int a = 0; a < G; ++a)
        for (int b = 0; b < G; ++b) {
            double h = hess[(size_t)a * G + b];
            if (!dense && h == 0.0) continue;
            for (int r = 0; r < R; ++r) tmp[(size_t)a * R + r] += h * J[(size_t)b * R + r];
        }
    for (int a = 0; a < G; ++a) {
        if (!dense && grad[a] == 0.0) {
            int any = 0;
            for (int r = 0; r < R; ++r) if (tmp[(size_t)a * R + r] != 0.0) { any = 1; break; }
            if (!any) continue;
        }
        for (int i = 0; i < R; ++i) {
            double ja = J[(size_t)a * R + i];
            if (dense || ja != 0.0)
                for (int r = 0; r < R; ++r) gamma[i * R + r] += 1e-8 * ja * tmp[(size_t)a * R + r];
        }
        double ga = grad[a];
        if (dense || ga != 0.0) {
            const double* Ca = C + (size_t)a * R * R;
            for (int e = 0; e < R * R; ++e) gamma[e] += 1e-8 * ga * Ca[e];
        }
    }
}

/*
 * Batch of vanilla OIS trades.  Schedules are shared INPUT tables (trade i uses schedule
 * sched[i]); every trade is valued independently.
 *   fixed leg of schedule s: entries fo[s]..fo[s+1] of f_pay_t, f_alpha
 *   float leg of schedule s: entries lo[s]..lo[s+1] of l_start_t, l_end_t, l_pay_t, l_alpha
 *   trade i: coupon, notional, spread, fixed_sign (+1 receive fixed, -1 pay fixed)
 * want: bit0 value, bit1 delta, bit2 gamma.  Outputs: pv[n], delta[n*R], gamma[n*R*R].
 */
void oracle_ois_batch_legs(int G, int R, int method, const double* x, const double* d, const double* J, const double* C,
                           const int64_t* fo, const double* f_pay_t, const double* f_alpha,
                           const int64_t* lo, const double* l_start_t, const double* l_end_t, const double* l_pay_t,
                           const double* l_alpha,
                           int64_t n, const int32_t* sched, const double* coupon, const double* notional,
                           const double* spread, const double* fixed_sign,
                           int want, int dense, int n_threads, int legs /* bit0 fixed, bit1 floating */,
                           double* pv, double* delta, double* gamma)
{
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        double* grad = (double*)malloc(sizeof(double) * G);
        double* hess = (want & 4) ? (double*)malloc(sizeof(double) * G * G) : NULL;
        double* tmp = (want & 4) ? (double*)malloc(sizeof(double) * G * R) : NULL;
        const double one = 1.0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < n; ++i) {
            const int s = sched[i];
            double* dl = (want & 2) ? delta + i * R : NULL;
            double* gm = (want & 4) ? gamma + i * (int64_t)R * R : NULL;
            if (dl) memset(dl, 0, sizeof(double) * R);
            if (gm) memset(gm, 0, sizeof(double) * R * R);
            double value = 0.0;
            /* ---- fixed leg: sign * sum_{t > 0} alpha N c DF(t)  (value_time = 0, DF(0) = d[0] = 1) */
            if (legs & 1) {
                if (want & 6) memset(grad, 0, sizeof(double) * G);
                if (hess) memset(hess, 0, sizeof(double) * G * G);
                double leg = 0.0;
                for (int64_t k = fo[s]; k < fo[s + 1]; ++k) {
                    if (!(f_pay_t[k] > 0.0)) continue;
                    bracket_t b; interp(f_pay_t[k], x, G, method, &b);
                    double pay = f_alpha[k] * notional[i] * coupon[i];
                    leg += add_term(fixed_sign[i] * pay, &b, &one, 1, d, G, (want & 6) ? grad : NULL, hess);
                }
                value += leg;
                if (want & 6) chain(grad, hess, J, C, G, R, dense, dl, gm, tmp);
            }
            /* ---- floating leg: -sign * sum_{p >= 0} ((DF(s)/DF(e) - 1)/a + spread) a N DF(p) */
            if (legs & 2) {
                if (want & 6) memset(grad, 0, sizeof(double) * G);
                if (hess) memset(hess, 0, sizeof(double) * G * G);
                double leg = 0.0;
                const double sg = -fixed_sign[i];
                for (int64_t k = lo[s]; k < lo[s + 1]; ++k) {
                    if (!(l_pay_t[k] >= 0.0)) continue;
                    bracket_t b[3]; double ex[3] = {1.0, -1.0, 1.0};
                    interp(l_start_t[k], x, G, method, &b[0]);
                    interp(l_end_t[k], x, G, method, &b[1]);
                    interp(l_pay_t[k], x, G, method, &b[2]);
                    double* gp = (want & 6) ? grad : NULL;
                    if (l_alpha[k] > 0.0) {
                        leg += add_term(sg * notional[i], b, ex, 3, d, G, gp, hess);          /* N DF_s DF_p / DF_e */
                        leg += add_term(-sg * notional[i], &b[2], &one, 1, d, G, gp, hess);   /* - N DF_p */
                    }
                    if (spread[i] != 0.0)
                        leg += add_term(sg * spread[i] * l_alpha[k] * notional[i], &b[2], &one, 1, d, G, gp, hess);
                }
                value += leg;
                if (want & 6) chain(grad, hess, J, C, G, R, dense, dl, gm, tmp);
            }
            if (want & 1) pv[i] = value;
        }
        free(grad); free(hess); free(tmp);
    }
}

void oracle_ois_batch(int G, int R, int method, const double* x, const double* d, const double* J, const double* C,
                      const int64_t* fo, const double* f_pay_t, const double* f_alpha,
                      const int64_t* lo, const double* l_start_t, const double* l_end_t, const double* l_pay_t,
                      const double* l_alpha,
                      int64_t n, const int32_t* sched, const double* coupon, const double* notional,
                      const double* spread, const double* fixed_sign,
                      int want, int dense, int n_threads, double* pv, double* delta, double* gamma)
{
    oracle_ois_batch_legs(G, R, method, x, d, J, C, fo, f_pay_t, f_alpha, lo, l_start_t, l_end_t, l_pay_t, l_alpha, n, sched,
                          coupon, notional, spread, fixed_sign, want, dense, n_threads, 3, pv, delta, gamma);
}

/*
 * CPU baseline WITH the unit factorisation the GPU path uses (bench.py cpu_baseline.value_units_port): a vanilla OIS is
 * linear in (coupon x notional, notional) once its schedule is fixed, so the legs of every distinct schedule are valued
 * once (oracle_ois_batch_legs on unit trades) and every trade is a weighted sum of two unit rows.  This is the expansion:
 *   out[i] = wa[i] * unit[ua[i]] + wf[i] * unit[uf[i]],   rows of 1 (pv) / R (delta) / R*R (gamma) doubles.
 */
void oracle_expand_units(int64_t n, int R, const int32_t* ua, const int32_t* uf, const double* wa, const double* wf,
                         const double* u_pv, const double* u_delta, const double* u_gamma, int n_threads,
                         double* pv, double* delta, double* gamma)
{
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    const int RR = R * R;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const double a = wa[i], f = wf[i];
        const int64_t x = ua[i], y = uf[i];
        if (pv) pv[i] = a * u_pv[x] + f * u_pv[y];
        if (delta) for (int r = 0; r < R; ++r) delta[i * R + r] = a * u_delta[x * R + r] + f * u_delta[y * R + r];
        if (gamma) for (int e = 0; e < RR; ++e) gamma[i * RR + e] = a * u_gamma[x * RR + e] + f * u_gamma[y * RR + e];
    }
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
