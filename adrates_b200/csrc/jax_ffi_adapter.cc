// jax_ffi_adapter.cc - XLA FFI handlers over the C ABI (include/adrates_b200.h), so that the valuation is a jax.ffi custom
// call inside jitted JAX programs and jax.grad / jax.hessian compose through it (adrates_b200/jax_binding.py wraps the call
// in custom_jvp rules: first order = the delta ladder, second order = the gamma matrix).
//
// Replaces what the reference does with JAX transforms around its own leg pricers:
//     grad(lambda d: price(d))(dfs) / hessian(...)(dfs) + the chain rule through Engine._cached_curve
//     cavour/market/position/engine.py:2551-2568 (fixed leg), :2909-2926 (floating leg)
//
// Built only where jaxlib's headers are present (adrates_b200/build.py::build_jax_ffi looks for xla/ffi/api/ffi.h under
// jaxlib/include); the image this repository is developed in has no JAX, see INTEGRATION.md section 3.
//
// Handler "cav_portfolio_totals":
//     operand   rates  f64[R]      par rates of the curve, in device memory (a traced value: d/d rates is what AD asks for)
//     attrs     ctx    i64         the cav_ctx* of a context whose curve plan and portfolio have been set up from Python
//               mask   i32         CAV_REQ_VALUE | CAV_REQ_DELTA | CAV_REQ_GAMMA
//     result    totals f64[1057]   [PV, ladder(32) per bp, gamma(32x32) per bp^2] of the whole portfolio
// Everything is enqueued on XLA's stream (cav_set_stream is a no-op once the context runs on it): the curve is re-bootstrapped
// from the device-resident rates (cav_curve_rebuild_dev), the portfolio revalued (cav_portfolio_value), no host round trip.
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

#include "xla/ffi/api/ffi.h"

#include "../../include/adrates_b200.h"

namespace ffi = xla::ffi;

static ffi::Error CavPortfolioTotalsImpl(cudaStream_t stream, int64_t ctx_handle, int32_t mask, ffi::Buffer<ffi::F64> rates,
                                         ffi::ResultBuffer<ffi::F64> totals) {
    cav_ctx* ctx = reinterpret_cast<cav_ctx*>(static_cast<intptr_t>(ctx_handle));
    if (!ctx) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "cav_portfolio_totals: null context");
    if (totals->element_count() != 1 + CAV_R + CAV_R * CAV_R)
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "cav_portfolio_totals: result must be f64[1057]");
    if (rates.element_count() < 1 || rates.element_count() > CAV_R)
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "cav_portfolio_totals: 1..32 par rates");
    int rc = cav_set_stream(ctx, static_cast<void*>(stream));
    if (rc == CAV_OK) rc = cav_curve_rebuild_dev(ctx, rates.typed_data());
    if (rc == CAV_OK) rc = cav_portfolio_value(ctx, static_cast<uint32_t>(mask), nullptr, nullptr, nullptr, totals->typed_data());
    if (rc != CAV_OK) return ffi::Error(ffi::ErrorCode::kInternal, std::string("adrates_b200: ") + cav_last_error(ctx));
    return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(CavPortfolioTotals, CavPortfolioTotalsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ctx")
                                  .Attr<int32_t>("mask")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
