// TEMPORARY placeholder until the batched DLS-PnP RANSAC lands (next commit): every entry point
// fails loudly.
#include "common.cuh"
struct cb_pnp { int dummy; };
extern "C" {
int cb_pnp_create(cb_pnp** out, int, int, int, int) { if (out) *out = nullptr; return cb::fail(CB_EINVAL, "pnp not built yet"); }
int cb_pnp_destroy(cb_pnp*) { return CB_OK; }
int cb_pnp_solve_batch(cb_pnp*, int, const int32_t*, const double*, const double*, const cb_ransac_params*, const int32_t*, double*, float*, int32_t*, int32_t*, int32_t*) { return cb::fail(CB_EINVAL, "pnp not built yet"); }
int cb_pnp_solve_batch_device(cb_pnp*, int, const int32_t*, int, const double*, const double*, const cb_ransac_params*, const int32_t*, double*, float*, int32_t*, int32_t*, int32_t*, void*) { return cb::fail(CB_EINVAL, "pnp not built yet"); }
int cb_pnp_dls_minimal(cb_pnp*, int, int, const double*, const double*, int32_t*, double*, double*) { return cb::fail(CB_EINVAL, "pnp not built yet"); }
}
