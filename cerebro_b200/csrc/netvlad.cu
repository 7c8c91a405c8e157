// TEMPORARY placeholder until the NetVLAD forward lands: every entry point fails loudly.
#include "common.cuh"
struct cb_descriptor { int dummy; };
extern "C" {
int cb_descriptor_create(cb_descriptor** out, const cb_netvlad_weights*, int, int, int, int, int) { if (out) *out = nullptr; return cb::fail(CB_EINVAL, "descriptor not built yet"); }
int cb_descriptor_destroy(cb_descriptor*) { return CB_OK; }
int cb_descriptor_dim(const cb_descriptor*) { return -1; }
int cb_descriptor_compute(cb_descriptor*, int, const uint8_t*, int64_t, float*) { return cb::fail(CB_EINVAL, "descriptor not built yet"); }
int cb_descriptor_compute_device(cb_descriptor*, int, const uint8_t*, float*, void*) { return cb::fail(CB_EINVAL, "descriptor not built yet"); }
int64_t cb_descriptor_get_activation(cb_descriptor*, int, float*, int64_t) { return cb::fail(CB_EINVAL, "descriptor not built yet"); }
}
