// placeholder replaced by the real kernels
#include "fo_internal.h"
#define STUB(name, ...) extern "C" int name(__VA_ARGS__) { return FO_ERR_UNSUPPORTED; }
STUB(fo_sph_isoft_argmax, fo_ctx*, const double*, int64_t, int64_t, int, int64_t*, double*, double*, double*)
STUB(fo_sph_coeffs_direct, fo_ctx*, const double*, const double*, int64_t, int64_t, int64_t, double, double*, int32_t*)
STUB(fo_sph_align_pairs, fo_ctx*, const double*, const double*, int64_t, int64_t, int64_t, double, int, int64_t*, double*, double*, double*, int32_t*)
STUB(fo_sph_align_pairs_dev, fo_ctx*, const double*, const double*, int64_t, int64_t, int64_t, double, int, int64_t*, double*, double*, double*, int32_t*)
STUB(fo_sph_harm_coeffs, fo_ctx*, const double*, int64_t, int64_t, int64_t, int64_t, double, double, double*, int32_t*)
STUB(fo_sph_bank_create, fo_ctx*, const double*, int64_t, int64_t, int64_t, int64_t, double, double, fo_bank**)
STUB(fo_sph_align_bank, fo_ctx*, const fo_bank*, const int64_t*, int64_t, int, int64_t*, double*, double*, double*, double*)
STUB(fo_sph_wigner_table, fo_ctx*, int64_t, double*)
