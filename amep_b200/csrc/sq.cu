// sq.cu -- static structure factor kernels (placeholder until the S(q) kernels land).
#include "amepb_internal.cuh"

using namespace amepb;

extern "C" {

int amepb_sq_harmonics(amepb_ctx* ctx, const void*, int, int, int64_t, int64_t, const double*, int64_t,
                       int32_t, int32_t, double*, void*) {
    return fail(ctx, AMEPB_E_INVAL, "amepb_sq_harmonics: not built yet");
}

int amepb_sf2d_modes(amepb_ctx* ctx, const void*, int, int, int64_t, int64_t, const double*, int64_t, int32_t,
                     int, int64_t, void*, void*) {
    return fail(ctx, AMEPB_E_INVAL, "amepb_sf2d_modes: not built yet");
}

}  // extern "C"
