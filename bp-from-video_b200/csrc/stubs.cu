// temporary: entry points not implemented yet
#include "common.cuh"
#define STUB(name, ...) extern "C" int name(__VA_ARGS__) { bpv::set_error(#name ": not built yet"); return BPV_E_INVALID; }
STUB(bpv_window_preprocess, const double*, const double*, const bpv_window_params*, double*, double*, int32_t*, void*)
STUB(bpv_window_spectrum, const double*, const double*, const bpv_window_params*, int32_t, float*, float*, int32_t*, int32_t*, double*, double*, void*)
STUB(bpv_window_xcorr, const double*, const double*, const bpv_window_params*, float*, float*, int32_t*, int32_t*, double*, double*, void*)
STUB(bpv_butter_sos_design, const double*, int32_t, const bpv_window_params*, double*, void*)
STUB(bpv_firls_design, const double*, int32_t, const bpv_window_params*, double*, void*)
