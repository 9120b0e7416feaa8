// cav_book_core.h - the rules of the device-side book flattener, written once as host/device inline functions:
// date arithmetic on int64 day serials, business-day roll, schedule generation, day-count fractions, the term lists of a
// vanilla OIS schedule class, bracket planning against the engine grid and the K rows of a tile group.
//
// The CUDA kernels in cav_book.cu call these per trade / per schedule class / per term / per signature group; the CPU
// test suite compiles the same header with g++ (tests/native/book_core_host.cpp) and compares it with the numpy restatement
// adrates_b200/batch.py, which is itself pinned by the reference's 1750 schedules and 300 day-count rows.
//
// What each block replaces in the reference (per trade, as Python objects):
//   cavour/utils/date.py:529-653, 796-879     Date.add_months / add_tenor
//   cavour/utils/calendar.py:139-217          Calendar.adjust (WEEKEND / NONE in closed form; holiday calendars by a walk over the
//                                             non-business-day bitmap of adrates_b200/holidays.py, calendar.py:257-1099)
//   cavour/utils/schedule.py:163-270          Schedule.generate (backward / forward roll, termination adjust, duplicate filter)
//   cavour/utils/day_count.py:122-330         DayCount.year_frac (two-date conventions)
//   cavour/utils/helpers.py:154-197           times_from_dates
//   cavour/market/position/engine.py:2519-2539, 2858-2897   per-leg host preparation (live-cashflow masks, leg signs)
//   cavour/market/curves/interpolator_ad.py:210-243         bracket rules of InterpolatorAd.simple_interpolate
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define CAVB_HD __host__ __device__ __forceinline__
#else
#define CAVB_HD inline
#endif

#if !defined(__CUDACC__)
struct int2 { int x, y; };
#endif

namespace cavb {

// enumerations mirror adrates_b200/dates.py (= cavour/utils/day_count.py, calendar.py)
enum { DC_ZERO = 0, DC_30_360_BOND = 1, DC_30E_360 = 2, DC_30E_360_ISDA = 3, DC_30E_PLUS_360 = 4, DC_ACT_ACT_ISDA = 5,
       DC_ACT_ACT_ICMA = 6, DC_ACT_365F = 7, DC_ACT_360 = 8, DC_ACT_365L = 9, DC_SIMPLE = 10 };
enum { BD_NONE = 1, BD_FOLLOWING = 2, BD_MOD_FOLLOWING = 3, BD_PRECEDING = 4, BD_MOD_PRECEDING = 5 };
enum { CAL_NONE = 1, CAL_WEEKEND = 2, CAL_LAST = 16 };   // 3..16: holiday calendars, need a bitmap (CalRef::bits)
enum { DG_FORWARD = 1, DG_BACKWARD = 2 };

// error bits (OR-ed into a device flag word; the host turns them into the reference's LibError messages)
enum { E_START_AFTER_MAT = 1, E_EFF_GE_TERM = 2, E_NOT_MONOTONIC = 4, E_SHORT_SCHEDULE = 8, E_TIME_ORDER = 16,
       E_TOO_MANY_DATES = 32, E_SIG_COLLISION = 64, E_KEY_RANGE = 128, E_CAL_RANGE = 256 };

// A calendar as the date rules see it: its CalendarTypes value and, for the holiday calendars, the bitmap of non-business
// days (weekend or holiday; bit (i & 31) of word (i >> 5) = day serial base + i, i < ndays).  An int converts implicitly
// (WEEKEND / NONE need no table), so call sites that pass a CalendarTypes value read as before.
struct CalRef {
    int type;
    int base, ndays;
    const uint32_t* bits;
    CAVB_HD CalRef(int t = CAL_WEEKEND) : type(t), base(0), ndays(0), bits(nullptr) {}
    CAVB_HD CalRef(int t, const uint32_t* b, int base_, int nd) : type(t), base(base_), ndays(nd), bits(b) {}
    CAVB_HD bool closed(int n) const {              // not a business day; outside the table only weekends are known
        const int i = n - base;
        if (i < 0 || i >= ndays) { int w = (n + 2) % 7; if (w < 0) w += 7; return w >= 5; }
        return (bits[i >> 5] >> (i & 31)) & 1u;
    }
    CAVB_HD bool covers(int64_t lo, int64_t hi) const { return lo >= base && hi < (int64_t)base + ndays; }
};

struct Conv {                 // book-wide conventions (per currency in practice)
    int64_t value_dt;         // day serial of the curve's value date
    int fixed_step, float_step;   // months per coupon period (12 / annual_frequency)
    int fixed_dc, float_dc;       // DayCountTypes values
    CalRef cal;                   // CalendarTypes value (+ the non-business-day bitmap of a holiday calendar)
    int bd, dg;                   // BusDayAdjustTypes / DateGenRuleTypes values
    int eom;                      // end-of-month roll
};

// Day serials, years and month counts fit 32 bits with room to spare (a serial is < 2^30 until the year 2.9 million; the
// entry point rejects anything larger): the date arithmetic runs in int, which matters on the GPU where 64-bit integer
// division is an order of magnitude slower.
CAVB_HD int fdiv(int a, int b) {                      // Python's floor division
    const int q = a / b;
    return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q;
}

CAVB_HD void ymd(int64_t n64, int& d, int& m, int64_t& y) {
    const int n = (int)n64;
    const int era = fdiv(n, 146097);
    const int doe = n - era * 146097;
    const int yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const int doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const int mp = (5 * doy + 2) / 153;
    d = doy - (153 * mp + 2) / 5 + 1;
    m = mp < 10 ? mp + 3 : mp - 9;
    y = yoe + era * 400 + (m <= 2);
}

CAVB_HD int64_t ordinal(int d, int m, int64_t y64) {
    const int yy = (int)y64 - (m <= 2);
    const int era = fdiv(yy, 400);
    const int yoe = yy - era * 400;
    const int mp = (m + 9) % 12;
    const int doy = (153 * mp + 2) / 5 + d - 1;
    return (int64_t)(era * 146097 + yoe * 365 + yoe / 4 - yoe / 100 + doy);
}

CAVB_HD bool is_leap(int64_t y64) { const int y = (int)y64; return ((y % 4 == 0) && (y % 100 != 0)) || (y % 400 == 0); }

CAVB_HD int days_in_month(int m, int64_t y) {
    const int base = (m == 2) ? 28 : ((m == 4 || m == 6 || m == 9 || m == 11) ? 30 : 31);
    return base + ((m == 2) && is_leap(y));
}

CAVB_HD int weekday(int64_t n) {                     // 0 = Monday (0000-03-01 is a Wednesday)
    const int w = ((int)n + 2) % 7;
    return w < 0 ? w + 7 : w;
}

// Date.add_months: same day of month (or `day` >= 1), clipped to the month length; eom -> month end
CAVB_HD int64_t add_months(int64_t n, int64_t mm, bool eom, int day = -1) {
    int d, m;
    int64_t y;
    ymd(n, d, m, y);
    const int k = (int)y * 12 + (m - 1) + (int)mm;
    const int y2 = fdiv(k, 12);
    const int m2 = k - y2 * 12 + 1;
    const int dim = days_in_month(m2, y2);
    const int want = day < 0 ? d : day;
    const int d2 = eom ? dim : (want < dim ? want : dim);
    return ordinal(d2, m2, y2);
}

// Date.add_tenor for 'Y' / 'M' counts: year tenors are applied one year at a time by the reference, so a 29-Feb start
// drops to the 28th and stays there; month tenors keep the original day
CAVB_HD int64_t add_tenor(int64_t n, int64_t count, bool years) {
    if (!years) return add_months(n, count, false);
    int d, m;
    int64_t y;
    ymd(n, d, m, y);
    const int day = (m == 2 && d == 29 && count != 0) ? 28 : d;
    return add_months(n, 12 * count, false, day);
}

// Calendar.adjust: weekday arithmetic for the WEEKEND / NONE calendars, a walk over the bitmap for the holiday calendars
CAVB_HD int64_t adjust(int64_t n, int bd, const CalRef& cal) {
    if (cal.type == CAL_NONE || bd == BD_NONE) return n;
    const bool following = (bd == BD_FOLLOWING || bd == BD_MOD_FOLLOWING);
    const bool modified = (bd == BD_MOD_FOLLOWING || bd == BD_MOD_PRECEDING);
    if (cal.bits) {
        const int step = following ? 1 : -1;
        int out = (int)n;
        while (cal.closed(out)) out += step;
        if (modified && out != (int)n) {
            int d0, m0, d1, m1;
            int64_t y0, y1;
            ymd(n, d0, m0, y0);
            ymd(out, d1, m1, y1);
            if (m1 != m0) { out = (int)n; while (cal.closed(out)) out -= step; }
        }
        return out;
    }
    const int w = weekday(n);
    const int fwd = w == 5 ? 2 : (w == 6 ? 1 : 0);
    const int bwd = -(w == 5 ? 1 : (w == 6 ? 2 : 0));
    const int first = following ? fwd : bwd, other = following ? bwd : fwd;
    int64_t out = n + first;
    if (modified) {
        int d0, m0, d1, m1;
        int64_t y0, y1;
        ymd(n, d0, m0, y0);
        ymd(out, d1, m1, y1);
        if (m1 != m0) out = n + other;
    }
    return out;
}

CAVB_HD bool dc_supported(int dc) {
    return dc == DC_ACT_365F || dc == DC_ACT_360 || dc == DC_SIMPLE || dc == DC_30_360_BOND || dc == DC_30E_360 ||
           dc == DC_30E_360_ISDA || dc == DC_30E_PLUS_360 || dc == DC_ACT_ACT_ISDA || dc == DC_ZERO;
}

// DayCount(dc).year_frac(dt1, dt2)[0] for the conventions that need only the two dates
CAVB_HD double year_frac(int64_t n1, int64_t n2, int dc) {
    if (dc == DC_ACT_365F || dc == DC_SIMPLE) return (double)(n2 - n1) / 365.0;
    if (dc == DC_ACT_360) return (double)(n2 - n1) / 360.0;
    int d1, m1, d2, m2;
    int64_t y1, y2;
    ymd(n1, d1, m1, y1);
    ymd(n2, d2, m2, y2);
    if (dc == DC_30_360_BOND || dc == DC_30E_360 || dc == DC_30E_360_ISDA || dc == DC_30E_PLUS_360) {
        const bool feb1 = (m1 == 2) && (d1 == days_in_month(m1, y1));
        const bool feb2 = (m2 == 2) && (d2 == days_in_month(m2, y2));
        if (d1 == 31) d1 = 30;
        if (dc == DC_30_360_BOND) { if (d2 == 31 && d1 == 30) d2 = 30; }
        else if (dc == DC_30E_360) { if (d2 == 31) d2 = 30; }
        else if (dc == DC_30E_360_ISDA) { if (feb1) d1 = 30; if (d2 == 31 || feb2) d2 = 30; }
        else { if (d2 == 31) { m2 += 1; d2 = 1; } }
        return (double)(360 * (int)(y2 - y1) + 30 * (m2 - m1) + (d2 - d1)) / 360.0;
    }
    // ACT/ACT ISDA (and ZERO, which the reference routes the same way)
    const double den1 = is_leap(y1) ? 366.0 : 365.0, den2 = is_leap(y2) ? 366.0 : 365.0;
    if (y1 == y2) return (double)(n2 - n1) / den1;
    const int64_t k1 = ordinal(1, 1, y1 + 1) - n1;
    const int64_t k2 = n2 - ordinal(1, 1, y2);
    const double a = (double)k1 / den1, b = (double)k2 / den2;
    const double s = a + b;
    return s + ((double)(y2 - y1) - 1.0);
}

// ------------------------------------------------------------------------------------------------------------------
// Schedule(eff, term, freq, cal, bd, dg, adjust_termination_dt=True, end_of_month)._adjusted_dts as a function of the
// position: a class is described by (cnt, dup) and every date is recomputed from its index, so no per-class date array
// is kept (schedule.py:163-270 via batch.roll_schedules).
// ------------------------------------------------------------------------------------------------------------------
struct Sched {
    int64_t eff, term;
    int step;
    CalRef cal;
    int bd, dg, eom;
    int cnt;          // dates before the duplicate filter = cnt + 1
    int dup;          // head dates dropped by the duplicate filter
    int err;          // E_* bits
    const int32_t* tab;   // optional: the cnt + 1 raw dates, precomputed (device flattener: k_bk_sched_table); null = recompute
    CAVB_HD int n_dates() const { return cnt + 1 - dup; }
};

CAVB_HD int64_t sched_raw_date(const Sched& s, int pos) {     // position before the duplicate filter
    if (s.tab) return (int64_t)s.tab[pos];
    int64_t dt;
    if (s.dg == DG_BACKWARD) {
        const int k = s.cnt - pos;
        dt = (k == 0) ? s.term : add_months(s.term, -(int64_t)s.step * k, s.eom != 0);
        if (pos > 0 && k > 0) dt = adjust(dt, s.bd, s.cal);
    } else {
        const bool last = pos == s.cnt;
        dt = last ? s.term : adjust(add_months(s.eff, (int64_t)s.step * pos, false), s.bd, s.cal);
    }
    if (pos == 0 && dt < s.eff) dt = s.eff;
    if (pos == s.cnt) dt = adjust(s.term, s.bd, s.cal);        // adjust_termination_dt
    return dt;
}

CAVB_HD int64_t sched_date(const Sched& s, int pos) { return sched_raw_date(s, pos + s.dup); }

// a schedule whose (cnt, dup) an earlier make_sched has already derived and verified
CAVB_HD Sched sched_from(int64_t eff, int64_t term, int step, const CalRef& cal, int bd, int dg, int eom, int cnt, int dup) {
    Sched s;
    s.eff = eff; s.term = term; s.step = step; s.cal = cal; s.bd = bd; s.dg = dg; s.eom = eom;
    s.cnt = cnt; s.dup = dup; s.err = 0; s.tab = nullptr;
    return s;
}

// number of dates before the duplicate filter (cnt + 1) of a schedule; sets s.cnt or s.err
CAVB_HD void sched_header(Sched& s, int max_dates) {
    const int64_t eff = s.eff, term = s.term;
    const int step = s.step, dg = s.dg, eom = s.eom;
    if (eff >= term) { s.err = E_EFF_GE_TERM; return; }
    int de, me, dt, mt;
    int64_t ye, yt;
    ymd(eff, de, me, ye);
    ymd(term, dt, mt, yt);
    const int M = ((int)yt * 12 + mt) - ((int)ye * 12 + me);
    const int q = fdiv(M, step), r = M - q * step;
    int cnt = q + (r != 0);
    if (dg == DG_BACKWARD) {
        const int64_t same_month = (q == 0) ? term : add_months(term, -(int64_t)step * q, eom != 0);
        cnt += (r == 0) && (same_month > eff);
    } else {
        const int64_t same_month = add_months(eff, (int64_t)step * q, false);
        cnt += (r == 0) && (same_month < term);
    }
    if (cnt + 1 > max_dates) { s.err = E_TOO_MANY_DATES; return; }
    s.cnt = cnt;
}

CAVB_HD Sched make_sched(int64_t eff, int64_t term, int step, const CalRef& cal, int bd, int dg, int eom, int max_dates) {
    Sched s;
    s.eff = eff; s.term = term; s.step = step; s.cal = cal; s.bd = bd; s.dg = dg; s.eom = eom;
    s.cnt = 0; s.dup = 0; s.err = 0; s.tab = nullptr;
    sched_header(s, max_dates);
    if (s.err) return s;
    int64_t prev = sched_raw_date(s, 0);
    for (int pos = 1; pos <= s.cnt; ++pos) {
        const int64_t cur = sched_raw_date(s, pos);
        if (cur < prev) s.err |= E_NOT_MONOTONIC;
        if (cur == prev) s.dup += 1;
        prev = cur;
    }
    return s;
}

// ------------------------------------------------------------------------------------------------------------------
// Terms of the units of one schedule class of vanilla OIS with payment lag 0 (batch.OISBook.flatten, shared units):
//   part 0  annuity          sum_{t_i > 0} alpha_i DF(t_i)                    fixed-leg periods (engine.py:2430)
//   part 1  floating         sum over coupons with t_pay >= 0 and alpha > 0 of DF(t_start) - DF(t_end)   (engine.py:2695)
//   part 2  spread annuity   sum_{t_pay >= 0} alpha_i DF(t_pay)               only for classes holding a trade with a spread
// Terms of a part that hit the same time are summed in period order and exact zeros dropped (batch._merge_terms); the
// times of consecutive periods are non-decreasing, so the merge is a run-length pass over the periods (verified: a
// decreasing time raises E_TIME_ORDER and the caller falls back to the host flattener).
// Sink: void term(int part, double t, double amt).
// ------------------------------------------------------------------------------------------------------------------
template <class Sink>
CAVB_HD void flush_run(Sink& sink, int part, bool& open, double t, double sum) {
    if (open && sum != 0.0) sink.term(part, t, sum);
    open = false;
}

// part 0: annuity on the fixed-leg schedule
template <class Sink>
CAVB_HD int walk_annuity(const Conv& cv, const Sched& fx, Sink& sink) {
    if (fx.n_dates() < 2) return E_SHORT_SCHEDULE;          // a leg needs at least one accrual period
    int err = 0;
    bool open = false;
    double rt = 0.0, rs = 0.0;
    const int np = fx.n_dates() - 1;
    int64_t s = sched_date(fx, 0);
    for (int i = 0; i < np; ++i) {
        const int64_t e = sched_date(fx, i + 1);
        const double t = year_frac(cv.value_dt, e, cv.fixed_dc);
        if (t > 0.0) {
            const double alpha = year_frac(s, e, cv.fixed_dc);
            if (open && t == rt) rs += alpha;
            else {
                if (open && t < rt) err |= E_TIME_ORDER;
                flush_run(sink, 0, open, rt, rs);
                open = true; rt = t; rs = alpha;
            }
        }
        s = e;
    }
    flush_run(sink, 0, open, rt, rs);
    return err;
}

// part 1: floating leg, +1 at the accrual start and -1 at the accrual end of every live coupon, merged by time (ties: starts
// before ends, as the stable sort of [starts | ends] orders them; the sums are sums of +-1, exact in any order)
template <class Sink>
CAVB_HD int walk_float(const Conv& cv, const Sched& fl, Sink& sink) {
    if (fl.n_dates() < 2) return E_SHORT_SCHEDULE;
    int err = 0;
    const int np = fl.n_dates() - 1;
    int i = 0, j = 0;                       // next start / next end candidate (period indices)
    bool open = false;
    double rt = 0.0, rs = 0.0;
    double ts_i = 0.0, te_j = 0.0, last_s = 0.0, last_e = 0.0;
    bool have_s = false, have_e = false, seen_s = false, seen_e = false;
    for (;;) {
        while (!have_s && i < np) {         // advance the start stream to the next live coupon
            const int64_t s = sched_date(fl, i), e = sched_date(fl, i + 1);
            const double tp = year_frac(cv.value_dt, e, cv.float_dc);
            if (tp >= 0.0 && year_frac(s, e, cv.float_dc) > 0.0) {
                ts_i = year_frac(cv.value_dt, s, cv.float_dc);
                if (seen_s && ts_i < last_s) err |= E_TIME_ORDER;
                last_s = ts_i; seen_s = true; have_s = true;
            }
            ++i;
        }
        while (!have_e && j < np) {
            const int64_t s = sched_date(fl, j), e = sched_date(fl, j + 1);
            const double tp = year_frac(cv.value_dt, e, cv.float_dc);
            if (tp >= 0.0 && year_frac(s, e, cv.float_dc) > 0.0) {
                te_j = tp;
                if (seen_e && te_j < last_e) err |= E_TIME_ORDER;
                last_e = te_j; seen_e = true; have_e = true;
            }
            ++j;
        }
        if (!have_s && !have_e) break;
        const bool take_s = have_s && (!have_e || ts_i <= te_j);
        const double t = take_s ? ts_i : te_j, a = take_s ? 1.0 : -1.0;
        if (take_s) have_s = false; else have_e = false;
        if (open && t == rt) rs += a;
        else { flush_run(sink, 1, open, rt, rs); open = true; rt = t; rs = a; }
    }
    flush_run(sink, 1, open, rt, rs);
    return err;
}

// part 2: spread annuity on the floating-leg schedule
template <class Sink>
CAVB_HD int walk_spread(const Conv& cv, const Sched& fl, Sink& sink) {
    if (fl.n_dates() < 2) return E_SHORT_SCHEDULE;
    int err = 0;
    bool open = false;
    double rt = 0.0, rs = 0.0;
    const int np = fl.n_dates() - 1;
    int64_t s = sched_date(fl, 0);
    for (int i = 0; i < np; ++i) {
        const int64_t e = sched_date(fl, i + 1);
        const double t = year_frac(cv.value_dt, e, cv.float_dc);
        if (t >= 0.0) {
            const double alpha = year_frac(s, e, cv.float_dc);
            if (open && t == rt) rs += alpha;
            else {
                if (open && t < rt) err |= E_TIME_ORDER;
                flush_run(sink, 2, open, rt, rs);
                open = true; rt = t; rs = alpha;
            }
        }
        s = e;
    }
    flush_run(sink, 2, open, rt, rs);
    return err;
}

template <class Sink>
CAVB_HD int walk_class(const Conv& cv, const Sched& fx, const Sched& fl, bool with_spread, Sink& sink) {
    if (fx.n_dates() < 2 || fl.n_dates() < 2) return E_SHORT_SCHEDULE;
    int err = walk_annuity(cv, fx, sink) | walk_float(cv, fl, sink);
    if (with_spread) err |= walk_spread(cv, fl, sink);
    return err;
}

// ------------------------------------------------------------------------------------------------------------------
// Bracket planning (curves.plan_queries = interpolator_ad.py:210-243): ln DF(t) = wa L[a] + wb L[b] on the engine grid
// x[0..G) (sorted, duplicates kept).  1e-10 snap to the FIRST nearest node, +1e-12 bracket shift, searchsorted(side='right')
// duplicate semantics, end clamping, max(x, 1e-15) in the zero-rate weights.  Slots with weight 0 point at node 0.
// ------------------------------------------------------------------------------------------------------------------
CAVB_HD int lower_bound(const double* x, int n, double t) {     // first index with x[i] >= t   (searchsorted 'left')
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid] < t) lo = mid + 1; else hi = mid; }
    return lo;
}
CAVB_HD int upper_bound_d(const double* x, int n, double t) {   // first index with x[i] > t    (searchsorted 'right')
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid] <= t) lo = mid + 1; else hi = mid; }
    return lo;
}

CAVB_HD void plan_query(double t, const double* x, int G, bool lzr, int& na, int& nb, double& wa, double& wb) {
    const double SNAP_TOL = 1e-10, BRACKET_EPS = 1e-12, TIME_FLOOR = 1e-15, FLAT = 4.930380657631324e-32;
    const int r = lower_bound(x, G, t);
    const int lo = r - 1 < 0 ? 0 : (r - 1 > G - 1 ? G - 1 : r - 1);
    const int hi = r > G - 1 ? G - 1 : r;
    const double d_lo = fabs(t - x[lo]), d_hi = fabs(t - x[hi]);
    int near = d_lo <= d_hi ? lo : hi;
    near = lower_bound(x, G, x[near]);
    const bool snap = (d_lo < d_hi ? d_lo : d_hi) < SNAP_TOL;
    const double ts = t + BRACKET_EPS;
    int b = upper_bound_d(x, G, ts);
    b = b < 1 ? 1 : (b > G - 1 ? G - 1 : b);
    int a = b - 1;
    const double dx = x[b] - x[a];
    const bool flat = fabs(dx) <= FLAT;
    const double w = flat ? 0.0 : (ts - x[a]) / (flat ? 1.0 : dx);
    const bool above = ts > x[G - 1], below = ts < x[0];
    double va, vb, va_hi, va_lo;
    if (lzr) {
        const double xa = x[a] > TIME_FLOOR ? x[a] : TIME_FLOOR, xb = x[b] > TIME_FLOOR ? x[b] : TIME_FLOOR;
        const double omw = 1.0 - w;
        const double tw = t * omw, tv = t * w;
        va = tw / xa;
        vb = tv / xb;
        va_hi = t / (x[G - 1] > TIME_FLOOR ? x[G - 1] : TIME_FLOOR);
        va_lo = t / (x[0] > TIME_FLOOR ? x[0] : TIME_FLOOR);
    } else {
        va = 1.0 - w; vb = w; va_hi = 1.0; va_lo = 1.0;
    }
    if (above) { a = G - 1; b = G - 1; va = va_hi; vb = 0.0; }
    else if (below) { a = 0; b = 0; va = va_lo; vb = 0.0; }
    if (snap) { a = near; b = near; va = 1.0; vb = 0.0; }
    na = (va == 0.0) ? 0 : a;
    nb = (vb == 0.0) ? 0 : b;
    wa = va; wb = vb;
}

// ------------------------------------------------------------------------------------------------------------------
// Tile-plan signature of a term (adrates_b200/tiles.py): kind 0 = grid snap (weights (1, 0)), 1 = one node with another
// weight (end clamp / zero second weight), 2 = bracket; units with equal key sequences share their K rows.
// ------------------------------------------------------------------------------------------------------------------
CAVB_HD int64_t term_key(double w0, double w1, int n0, int n1) {
    const int64_t kind = (w0 == 1.0 && w1 == 0.0) ? 0 : (w1 == 0.0 ? 1 : 2);
    const int64_t b = kind == 2 ? n1 : 0;
    return (kind << 40) | ((int64_t)n0 << 20) | b;
}

CAVB_HD uint64_t mix64(uint64_t h) {                   // splitmix64 finaliser
    h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 27; h *= 0x94D049BB133111EBull;
    h ^= h >> 31;
    return h;
}

// K rows of a signature group (tiles.plan_tiles): per term, in position order,
//   kind 0:  C_a  x p
//   kind 1:  H_a  x p w0,  G_aa x p w0^2
//   kind 2:  + H_b x p w1,  G_bb x p w1^2,  G_ab x p w0 w1
// A table row met again inside the same 32-position chunk by the LATEST K row of that table row is merged into it as a
// second contribution (k_pos2 / k_coef2) if that row has none yet.  Rows are handed to the sink in emission order:
//   int  find(int row, int pos)   index of the latest K row of this chunk with this table row whose second slot is free, or -1
//                                 if the latest K row with this table row lies in an earlier chunk / is full / does not exist
//   void merge(int k, int pos, int coef)
//   void emit(int row, int pos, int coef)
// The sinks below implement find() by a backward scan over the rows of the current chunk.
enum { COEF_P = 0, COEF_PW0 = 1, COEF_PW1 = 2, COEF_PW0SQ = 3, COEF_PW1SQ = 4, COEF_PW0W1 = 5 };

template <class Sink>
CAVB_HD void emit_row(Sink& sink, int row, int pos, int coef) {
    const int hit = sink.find(row, pos);
    if (hit >= 0) sink.merge(hit, pos, coef);
    else sink.emit(row, pos, coef);
}

template <class Sink>
CAVB_HD void emit_term_rows(Sink& sink, int64_t key, int pos, int G, const int* pair_index) {
    const int kind = (int)(key >> 40), na = (int)((key >> 20) & 0xFFFFF), nb = (int)(key & 0xFFFFF);
    if (kind == 0) emit_row(sink, G + na, pos, COEF_P);
    else {
        emit_row(sink, na, pos, COEF_PW0);
        emit_row(sink, 2 * G + na, pos, COEF_PW0SQ);
        if (kind == 2) {
            emit_row(sink, nb, pos, COEF_PW1);
            emit_row(sink, 2 * G + nb, pos, COEF_PW1SQ);
            emit_row(sink, 3 * G + (pair_index ? pair_index[na] : na), pos, COEF_PW0W1);
        }
    }
}

// K-row sink over packed rows (x = table row; y = pos | coef << 8 | pos2 << 16 | coef2 << 24, coef2 = 7: none), the format
// k_units_mma reads.  count-only when `pack` is null.
struct KRowSink {
    int2* pack;            // output (may be null: count only)
    int n;                 // rows emitted so far
    int chunk_first;       // index of the first row of the current 32-position chunk
    int chunk;             // current chunk id (pos >> 5)
    // count-only mode keeps the rows of the current chunk in a small ring (at most 160 per chunk)
    int ring_row[160];
    unsigned char ring_full[160];
    CAVB_HD KRowSink(int2* p) : pack(p), n(0), chunk_first(0), chunk(-1) {}
    CAVB_HD void roll(int pos) {
        if ((pos >> 5) != chunk) { chunk = pos >> 5; chunk_first = n; }
    }
    CAVB_HD int find(int row, int pos) {
        roll(pos);
        for (int k = n - 1; k >= chunk_first; --k)
            if (ring_row[k - chunk_first] == row) return ring_full[k - chunk_first] ? -1 : k;
        return -1;
    }
    CAVB_HD void merge(int k, int pos, int coef) {
        ring_full[k - chunk_first] = 1;
        if (pack) pack[k].y = (pack[k].y & 0xFFFF) | (pos << 16) | (coef << 24);
    }
    CAVB_HD void emit(int row, int pos, int coef) {
        roll(pos);
        const int s = n - chunk_first;
        if (s < 160) { ring_row[s] = row; ring_full[s] = 0; }
        if (pack) { pack[n].x = row; pack[n].y = pos | (coef << 8) | (7 << 24); }
        ++n;
    }
};

}  // namespace cavb
