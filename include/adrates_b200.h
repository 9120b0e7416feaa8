/*
 * adrates_b200.h - C ABI of the B200-native valuation-and-Greeks library.
 *
 * The reference (ludcode/ADRates, "Cavour") has no FFI: its boundary for this path is the
 * Python API Model.build_curve / curve.df_ad / Position.compute([VALUE, DELTA, GAMMA]) /
 * Portfolio.compute.  The entry points below are what a Python (ctypes / jax.ffi) binding
 * for that path calls; each one names the reference code it replaces.  See
 * INTEGRATION.md for the reference-side stubs.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, IEEE FP64 only; no exceptions cross the ABI.
 *   - every function returns CAV_OK (0) or a negative CAV_E_* code; cav_last_error(ctx)
 *     returns a message for the last failure on that context.
 *   - a context binds one CUDA device and owns one stream; calls on one context are
 *     serialised on that stream.  Contexts are independent (one per GPU / per thread).
 *   - pointer arguments are HOST pointers unless the name ends in `_dev`; caller owns
 *     every buffer it passes in.  Functions that take `_dev` outputs are asynchronous on
 *     the context stream; cav_sync() waits for them.
 *   - pillars: n_rates <= 32 (one warp lane per par-rate pillar); ladders and gamma rows
 *     are always written 32 wide, zero padded.
 */
#ifndef ADRATES_B200_H
#define ADRATES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAV_OK            0
#define CAV_E_INVALID    -1   /* bad argument */
#define CAV_E_CUDA       -2   /* CUDA runtime error (message has the detail) */
#define CAV_E_STATE      -3   /* call order: curve / portfolio not set */
#define CAV_E_UNSUPPORTED -4  /* e.g. n_rates > 32, unknown interpolation */

#define CAV_R 32              /* ladder width */

/* interpolation methods: InterpTypes values, cavour/utils/global_types.py:76-84 */
#define CAV_INTERP_FLAT_FWD_RATES    1
#define CAV_INTERP_LINEAR_ZERO_RATES 4

/* request mask: RequestTypes, cavour/utils/global_types.py:69-74 */
#define CAV_REQ_VALUE 1u
#define CAV_REQ_DELTA 2u
#define CAV_REQ_GAMMA 4u
/* multi-GPU: the portfolio totals the valuation returns are the sum over all ranks of cav_comm_init (see below) */
#define CAV_REQ_ALLREDUCE 8u

typedef struct cav_ctx cav_ctx;

/* ---- context --------------------------------------------------------------------- */
int         cav_create(cav_ctx** ctx, int device);
void        cav_destroy(cav_ctx* ctx);
const char* cav_last_error(const cav_ctx* ctx);
int         cav_version(void);
int         cav_sync(cav_ctx* ctx);
/* CUDA-event timer on the context's own stream (torch.cuda.Event cannot see it). */
int         cav_timer_start(cav_ctx* ctx);
int         cav_timer_stop(cav_ctx* ctx, float* elapsed_ms);   /* synchronises */
/* run this context's work on a caller-owned CUDA stream (e.g. the framework's current
 * stream, so its events and collectives order with the kernels here) */
int         cav_set_stream(cav_ctx* ctx, void* cuda_stream);
/* Pipelined host->device upload for the end-to-end call sequence upload -> [set_tiles] -> value.  The reference
 * rebuilds every Position's inputs on the host before it values it (position.py:55, engine.py:2519-2539); here
 * the per-trade arrays (comp_weight, out_index) travel on a side copy stream in group-aligned chunks while the
 * unit arrays, the tile plan and the units kernel proceed, and the expansion kernel of a chunk starts as soon as
 * its weights have landed (chunks alternate between two streams so that one chunk's last wave overlaps the next
 * chunk's first); neither cav_portfolio_upload nor cav_portfolio_set_tiles waits for its copies.  Host-side
 * validation still covers every index a kernel dereferences, but is split: cav_portfolio_upload checks what the
 * units kernel reads (node indices, unit offsets) and fails as usual; the per-trade arrays (group_offsets,
 * group_units, out_index) are checked by the next valuation call after it has launched the units kernel and
 * before it launches anything that reads them, or by cav_sync - that call then returns CAV_E_INVALID with the
 * same message and the portfolio is discarded.  Contract when enabled: every HOST buffer handed to
 * cav_portfolio_upload must stay valid and unmodified until the next cav_portfolio_value_host or cav_sync call on
 * this context has returned (the tile plan is staged through a pinned arena of the library and may be reused at
 * once).  Off by default.  Results are bit-identical either way. */
int         cav_set_async_upload(cav_ctx* ctx, int enable);
/* per-kernel CUDA-event timing of cav_portfolio_value: ms[0] units kernel, ms[1] per-trade
 * expansion kernel, ms[2] portfolio-total reduction (needs agg output) */
int         cav_profile(cav_ctx* ctx, int enable);
int         cav_last_kernel_ms(cav_ctx* ctx, float* ms);
/* number of kernels this context has launched since creation */
int64_t     cav_launch_count(const cav_ctx* ctx);

/* ---- curve: replaces Engine.build_curve_ad + Engine._cached_curve ------------------
 * (cavour/market/position/engine.py:2246-2412).  The host passes the rate-independent
 * plan of the engine grid (node times, coupon accruals, parent swap, previous-annuity
 * node: engine.py:2283-2334) and the par rates; the device runs the bootstrap recursion
 * (engine.py:2337-2349) together with its exact first/second-order tangents, giving
 * dfs[G], jac[G][R] (= jacrev, :2388) and hess[G][R][R] (= hessian, :2389), and derives
 * the log-DF tables the valuation kernels read.  order: 0 = dfs, 1 = +jac, 2 = +hess. */
int cav_curve_build(cav_ctx* ctx, int interp_method,
                    const double* swap_rates, int n_rates,
                    const double* node_time, const double* node_acc,
                    const int32_t* node_swap, const int32_t* node_prev, int n_nodes,
                    int order);
/* copy the bootstrapped tables back (any pointer may be NULL); jac is [G][n_rates],
 * hess is [G][n_rates][n_rates] - the shapes of the reference cache (engine.py:2405-2410) */
int cav_curve_read(cav_ctx* ctx, double* dfs, double* jac, double* hess);

/* Re-bootstrap the current plan from par rates that already live on the device (asynchronous on the
 * context stream): what a jax.ffi handler calls when the rates are a traced device buffer, and what a
 * scenario loop with full Greeks per scenario uses (Model.scenario, models.py:507-557). */
int cav_curve_rebuild_dev(cav_ctx* ctx, const double* swap_rates_dev);

/* Set the curve tables directly instead of bootstrapping them: dfs[G], jac[G][n_rates] (may be
 * NULL), hess[G][n_rates][n_rates] (may be NULL).  Used for curves whose bootstrap lives on the
 * host - XccyCurve._times/_dfs/_jac_basis (cavour/trades/rates/xccy_curve.py:529-703) - and for
 * the stacked (foreign OIS + XCCY) node grid of Engine._compute_xccy (engine.py:1411-1765). */
int cav_curve_set_tables(cav_ctx* ctx, const double* dfs, const double* jac, const double* hess, int n_nodes,
                         int n_rates);

/* ---- curve.df_ad: replaces DiscountCurve._linear_forward_interp -------------------
 * (cavour/market/curves/discount_curve.py:385-415) on the path-A nodes. */
int cav_df_ad(cav_ctx* ctx, const double* node_time, const double* node_df, int n_nodes,
              const double* t, int64_t n, double* out);

/* ---- XccyCurve bootstrap on the device: replaces XccyCurve._build_curve_ad and the jacrev / hessian through it
 * (cavour/trades/rates/xccy_curve.py:954-1206 recursion, :594 _jac_basis, :603-606 _hess_basis).
 * The payment points of the calibration swaps' foreign legs come from the host plan (xccy_curve.py:707-935), sorted by
 * (time, swap): pt_swap = pillar index, pt_flags = 1 notional exchange | 4 paid on the valuation date | 8 maturity of its
 * swap, pt_spread_sens = accrual x notional (0 for exchanges), pt_base = the cashflow at zero spread, pt_df_ois = foreign
 * OIS discount factor at the payment date, pt_pv_dom = PV of the swap's domestic leg (read at maturity points).
 * spreads[n_scen][n_spreads] are the pillar basis spreads of n_scen curves (base + shocked).  Outputs (host):
 * df_out[n_scen][n_points]; order >= 1: jac_out[n_scen][n_points][n_spreads] = d DF / d spread; order 2:
 * hess_out[n_scen][n_points][n_spreads][n_spreads].  n_spreads <= 32. */
int cav_xccy_curve_scan(cav_ctx* ctx, int n_points, int n_spreads, const double* pt_time, const int32_t* pt_swap,
                        const int32_t* pt_flags, const double* pt_spread_sens, const double* pt_base, const double* pt_df_ois,
                        const double* pt_pv_dom, double spot_fx, const double* spreads, int n_scen, int order,
                        double* df_out, double* jac_out, double* hess_out);

/* ---- portfolio: the flattened cashflow schedule ------------------------------------
 * Replaces the per-trade host prep of Engine._fixed_leg_analytics / _float_leg_analytics
 * (engine.py:2519-2539, 2858-2897).  A *unit* is a set of terms
 *       value = sum_i amt[i] * exp( sum_{m < n_pairs} weight[i][m] * L[node[i][m]] ),  L = ln dfs,
 * i.e. cashflows whose discount factors are the reference's interpolation of the engine
 * grid (interpolator_ad.py:210-243): each DF query is one bracket = two (node, weight)
 * pairs precomputed by the host planner (grid snap = weight 1 on one node).  n_pairs = 2
 * for cashflows discounted by a single DF; n_pairs = 6 lets a term be a product
 * DF(s)*DF(p)/DF(e) (floating coupons with a payment lag, engine.py:2675-2692).
 * A *trade* is a weighted sum of n_comp units (weight 0 = unused slot): a vanilla OIS is
 * {coupon*notional x annuity unit, notional x float unit}; an irregular trade is one
 * private unit with weight 1.  Trades are grouped so that the trades of a group share
 * their unit ids: group g covers trades [group_offsets[g], group_offsets[g+1]) (at most 256
 * trades; split longer runs) and uses units group_units[g*n_comp ..].  out_index (or NULL = identity) is the row of each
 * trade in the per-trade outputs.  unit_weight[n_units] (or NULL = computed here) is the
 * sum over trades of the weights on each unit; it turns unit results into portfolio totals. */
int cav_portfolio_upload(cav_ctx* ctx,
                         int64_t n_units, int64_t n_terms, const int64_t* unit_offsets,
                         int n_pairs, const double* amt, const double* weight, const int32_t* node,
                         int64_t n_trades, int n_comp, const double* comp_weight,
                         int64_t n_groups, const int64_t* group_offsets, const int32_t* group_units,
                         const int64_t* out_index, const double* unit_weight);

/* Optional tile plan for the tensor-core Greeks kernel (host planner: adrates_b200/tiles.py).  Units that
 * bracket the same node pairs term by term are grouped into tiles of 16; their gamma/delta rows are then one
 * FP64 GEMM per tile, [16 x K] coefficients times K rows of per-curve symmetric tables (H_n, C_n, g_n g_n^T and,
 * for the node pairs listed in `pairs`, g_a g_b^T + g_b g_a^T), evaluated with mma.sync.m8n8k4.f64.
 * tile_units[n_tiles][tile_size] (tile_size 16, -1 = padding) must cover every unit exactly once; tile t uses K rows
 * [tile_kstart[t], tile_kstart[t] + tile_kcount[t]) of (k_row = table row id, k_pos = term position within the
 * unit, k_coef = 0:p 1:p*w0 2:p*w1 3:p*w0^2 4:p*w1^2 5:p*w0*w1), ordered by k_pos, at most 160 per chunk of 32
 * positions.  k_pos2 / k_coef2 (both NULL, or k_coef2 = -1 per row for none): a second term of the same 32-position
 * chunk that feeds the same table row (its coefficient is added).  Table row ids: n (H), G+n (C), 2G+n (g g^T),
 * 3G+i (pair i).  Replaces the same reference code as cav_portfolio_value; it only changes how it is computed.
 * tile_mask[n_tiles] (or NULL = all pillars): bit q set when the pillar at position q of the permuted order can be
 * non-zero in the tile's Greeks (union of the supports of the nodes its terms touch;
 * adrates_b200/tiles.py::node_support_masks derives it from the bootstrap plan).  The tile GEMM then runs over the
 * packed columns of the active pillars only; the masks are checked on the device against the tables
 * (CAV_E_INVALID from the valuation call if a mask is too small).  Tiles must be ordered by size class
 * (compact columns na(na+3)/2 in steps of 64: <= 64, 128, 192, 256, 384, 576), one kernel instantiation per class.
 * pillar_perm[32] (or NULL = identity): position q of the packed tables holds par-rate pillar pillar_perm[q]; the
 * planner orders pillars by how much work uses them so that active sets are prefixes (contiguous table columns).
 * Must be called after every cav_portfolio_upload (which clears the plan). */
int cav_portfolio_set_tiles(cav_ctx* ctx, int n_tiles, int tile_size, const int32_t* tile_units, const int32_t* tile_kstart,
                            const int32_t* tile_kcount, int64_t n_krows, const int32_t* k_row, const int32_t* k_pos,
                            const int32_t* k_coef, const int32_t* k_pos2, const int32_t* k_coef2, int n_pair_rows,
                            const int32_t* pairs, const uint32_t* tile_mask, const int32_t* pillar_perm);

/* ---- explicit-cashflow PV on a path-A curve: replaces DiscountCurve.df / _df and Interpolator._uinterpolate
 * (cavour/market/curves/discount_curve.py:300-436, cavour/market/curves/interpolator.py:35-170) and the discounting
 * inside ZeroCouponInflationSwap.value / SwapInflationLeg.value (cavour/trades/rates/zcis.py:176-238,
 * cavour/trades/rates/swap_inflation_leg.py:166-236).
 * node_time[n_nodes] (scanned from the front like the reference, whose OIS curves carry duplicate node times) / node_df[n_nodes] are the
 * curve's `_times` / `_dfs`; interp_method is CAV_INTERP_FLAT_FWD_RATES or CAV_INTERP_LINEAR_ZERO_RATES.
 * cav_curve_df: out[i] = DF(t[i]); every t must be >= 0 (CAV_E_INVALID otherwise, like the reference's LibError).
 * cav_cashflow_pv: trade i owns cashflows [offsets[i], offsets[i+1]) at times t[] with signed amounts amt[];
 * pv[i] = sum amt * DF(t) / DF(t_value); *total = sum_i pv[i] (fixed-order reduction).  All pointers are host memory. */
int cav_curve_df(cav_ctx* ctx, int interp_method, const double* node_time, const double* node_df, int n_nodes,
                 const double* t, int64_t n, double* out);
int cav_cashflow_pv(cav_ctx* ctx, int interp_method, const double* node_time, const double* node_df, int n_nodes,
                    double t_value, int64_t n_trades, const int64_t* offsets, const double* t, const double* amt,
                    double* pv, double* total);
/* cav_cashflow_pv on DEVICE-resident cashflows: offsets_dev / t_dev / amt_dev / pv_dev / total_dev (may be NULL) are device
 * pointers, the call is asynchronous on the context stream and moves nothing but the curve nodes (host, <= 1024).  The times
 * are not scanned on the host: a cashflow with a negative time (the reference's LibError) makes its trade's PV NaN. */
int cav_cashflow_pv_dev(cav_ctx* ctx, int interp_method, const double* node_time, const double* node_df, int n_nodes,
                        double t_value, int64_t n_trades, const int64_t* offsets_dev, const double* t_dev, const double* amt_dev,
                        double* pv_dev, double* total_dev);

/* ---- valuation: replaces Position.compute / Portfolio.compute ----------------------
 * (cavour/market/position/position.py:62-80, cavour/market/portfolio/portfolio.py:39-67,
 * engine.py:153-189, 2541-2574, 2899-2932).
 * Per-trade outputs (device pointers, may be NULL): pv[n_trades], delta[n_trades][32]
 * (x1e-4, per bp), gamma[n_trades][32][32] (x1e-8, per bp^2, full symmetric matrix).
 * agg_dev (device, may be NULL): portfolio totals [1 + 32 + 1024] = sum over trades,
 * accumulated in a fixed order (bitwise reproducible for a given upload). */
int cav_portfolio_value(cav_ctx* ctx, uint32_t request_mask,
                        double* pv_dev, double* delta_dev, double* gamma_dev, double* agg_dev);
/* same, writing the portfolio totals to host memory (synchronises) */
int cav_portfolio_value_host(cav_ctx* ctx, uint32_t request_mask,
                             double* pv_dev, double* delta_dev, double* gamma_dev, double* agg_host);

/* ---- chain rule as a batched FP64 tensor-core GEMM -----------------------------------
 * Alternative delta path: replaces `jnp.dot(grad_dfs, jac) * 1e-4` (engine.py:2554-2555,
 * 2912-2913) for the whole book with delta[U][32] = Q[U][G] * (1e-4 J/d)[G][32] on the FP64
 * tensor pipe (mma.sync m8n8k4).  Writes per-trade pv[n_trades] (may be NULL) and
 * delta[n_trades][32]; returns the GEMM kernel's CUDA-event time and its flop count so the
 * caller can report tensor-pipe utilisation.  Private layouts must be uploaded in output
 * order (out_index == NULL). */
int cav_portfolio_delta_gemm(cav_ctx* ctx, double* pv_dev, double* delta_dev, float* gemm_ms, double* gemm_flops);

/* ---- scenarios: replaces Model.scenario + rebuild + Position.compute([VALUE]) ------
 * (cavour/models/models.py:507-557, engine.py:2337-2349).  Each row of shocked_rates is a
 * full par-rate vector; every curve is re-bootstrapped (DFs only) and every trade
 * revalued.  pnl_dev[s][trade] = PV under scenario s (device, [n_scen][n_trades]). */
int cav_scenarios(cav_ctx* ctx, const double* shocked_rates, int n_scen, double* pnl_dev);
/* what the scenario path found in the uploaded book (after a cav_scenarios call): out[4] = distinct DF queries, prefix chains
 * of units (runs of units whose term lists extend their predecessor's: same dates, growing maturity), terms the chain kernel
 * walks (against n_terms for the per-unit kernels), 1 if the last call took the chain kernel.  Measurement / tests only. */
int cav_scenarios_info(cav_ctx* ctx, int64_t* out);

/* ---- device-side book flattener: replaces the per-trade object layer in front of the valuation ------------------
 * The reference turns every OIS into a Schedule, two legs and ~50 Date objects before it values a cashflow
 * (cavour/utils/schedule.py:163-270, calendar.py:139-217, day_count.py:122-330, date.py:529-879, helpers.py:154-197,
 * cavour/trades/rates/swap_fixed_leg.py:130-196, swap_float_leg.py:130-186, engine.py:2519-2539, 2858-2897) and plans
 * every DF query at valuation time (cavour/market/curves/interpolator_ad.py:210-243).  cav_book_from_arrays takes a book
 * of vanilla OIS as per-trade ARRAYS and builds, on the device, exactly the layout cav_portfolio_upload +
 * cav_portfolio_set_tiles would have been given by the host flattener (adrates_b200/batch.py + tiles.py): dates are int64
 * day serials (days since 0000-03-01, adrates_b200.dates.Date._n); trades with equal (effective, termination) dates share
 * their schedule units (annuity, floating, spread annuity).  Afterwards cav_portfolio_value* / cav_scenarios work as after
 * an upload.  All pointers are HOST memory (pinned memory is copied without staging); the call returns when the book is
 * resident (the inputs may be reused).  Conventions are book-wide (per currency in practice).
 * maturity: `termination` (serials), or tenor[] counts in tenor_unit (Date.add_tenor 'nY' / 'nM' rules).
 * fixed_sign: +1 receive fixed / -1 pay fixed.  spread: floating spread per trade or NULL.
 * flags: CAV_BOOK_TILES also plans the tiles of the tensor-core Greeks kernel.
 * Errors: CAV_E_INVALID with the reference's LibError text ("Start date after maturity date", "Effective date must be
 * before termination date.", "Dates are not monotonic", "Schedule has none or only one date"); CAV_E_UNSUPPORTED for what
 * only the host flattener handles (payment lag, three-date day counts, a holiday calendar whose bitmap has not been set,
 * a fully matured book). */
#define CAV_TENOR_YEARS  0
#define CAV_TENOR_MONTHS 1
#define CAV_BOOK_TILES   1u
#define CAV_BOOK_DATES_I32 2u   /* `effective` / `termination` address int32 day serials (4 bytes per trade over the host link) */
#define CAV_BOOK_SIGN_I8   4u   /* `fixed_sign` addresses int8 values +1 (receive fixed) / -1 (pay fixed) */
typedef struct cav_book_conv {
    int64_t value_dt;            /* curve value date (serial) */
    int32_t fixed_freq_months;   /* 12 / annual_frequency(fixed_freq_type) */
    int32_t float_freq_months;
    int32_t fixed_dc, float_dc;  /* DayCountTypes values (cavour/utils/day_count.py:91-120) */
    int32_t cal_type;            /* CalendarTypes: 1 NONE, 2 WEEKEND; 3..16 holiday calendars (cav_book_set_holidays first) */
    int32_t bd_type;             /* BusDayAdjustTypes 1..5 */
    int32_t dg_type;             /* DateGenRuleTypes: 1 FORWARD, 2 BACKWARD */
    int32_t end_of_month;
    int32_t payment_lag;         /* must be 0 on this path */
} cav_book_conv;
/* Holiday calendar of the books flattened from now on.  Replaces the per-date if-chains of Calendar.is_holiday /
 * is_business_day (cavour/utils/calendar.py:257-1099) by a bitmap the schedule kernels walk in Calendar.adjust
 * (calendar.py:139-217): bit (i & 31) of word (i >> 5) is set when day serial base_serial + i is NOT a business day
 * (weekend or holiday), i < n_days (adrates_b200.holidays.table(cal).words(), 1901-2199).  Host pointer, copied before the
 * call returns; NULL clears it.  A book whose trades reach within two months of either end of the range is rejected
 * (CAV_E_INVALID). */
int cav_book_set_holidays(cav_ctx* ctx, const uint32_t* non_business_bits, int64_t base_serial, int64_t n_days);
int cav_book_from_arrays(cav_ctx* ctx, const cav_book_conv* conv, int64_t n_trades, const int64_t* effective,
                         const int64_t* termination, const int32_t* tenor, int tenor_unit, const double* fixed_sign,
                         const double* coupon, const double* notional, const double* spread, uint32_t flags);
/* sizes of the portfolio on the device: out[10] = n_units, n_terms, n_trades, n_groups, n_pairs, n_comp, n_tiles, n_krows,
 * n_pair_rows, built_on_device */
int cav_book_info(cav_ctx* ctx, int64_t* out);
/* copy the flat arrays / the tile plan of the portfolio on the device to the host (any pointer may be NULL; shapes as in
 * cav_portfolio_upload / cav_portfolio_set_tiles; k_desc = k_pos | k_coef << 8 | k_pos2 << 16 | k_coef2 << 24 with
 * k_coef2 = 7 for none; perm[32]; class_begin[7]) */
int cav_book_read(cav_ctx* ctx, int64_t* unit_offsets, double* amt, double* weight, int32_t* node, double* comp_weight,
                  int64_t* group_offsets, int32_t* group_units, int64_t* out_index, double* unit_weight);
int cav_book_read_tiles(cav_ctx* ctx, int32_t* tile_units, int32_t* tile_kstart, int32_t* tile_kcount, int32_t* tile_npos,
                        uint32_t* tile_mask, int32_t* k_row, int32_t* k_desc, int32_t* pairs, int32_t* perm, int32_t* class_begin);

/* ---- multi-GPU: portfolio totals over the GPUs of one box -------------------------------------------------------
 * Replaces the running sums of Portfolio.compute (cavour/market/portfolio/portfolio.py:48-65) when the book is sharded
 * over one process per GPU: the only exchange is the sum of the 1057 totals.  It is done inside the totals kernel over
 * peer memory (NVLink stores into symmetric buffers + sequence flags, csrc/cav_comm.cu), not by a collective library.
 * Setup: every rank calls cav_comm_local_handle, the ranks exchange the CAV_COMM_HANDLE_BYTES blobs by any means
 * (torch.distributed all_gather, MPI, a file), then every rank calls cav_comm_init with all blobs in rank order.
 * Afterwards cav_portfolio_value / cav_portfolio_value_host with CAV_REQ_ALLREDUCE in the request mask return the
 * whole-job totals on every rank (bit-identical across ranks); every rank must make the same sequence of such calls.
 * At most 8 ranks, one process per GPU (two ranks on one device are refused: a waiting kernel must never share a GPU
 * with the kernel it waits for).  A peer that never arrives yields NaN totals and lost != 0 in cav_comm_status. */
#define CAV_COMM_HANDLE_BYTES 96
int cav_comm_local_handle(cav_ctx* ctx, void* handle_out);
int cav_comm_init(cav_ctx* ctx, int rank, int world, const void* handles);
int cav_comm_status(cav_ctx* ctx, int* rank, int* world, int* lost);

#ifdef __cplusplus
}
#endif
#endif /* ADRATES_B200_H */
