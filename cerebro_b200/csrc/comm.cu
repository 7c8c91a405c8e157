// cb_comm_*: a communicator handle for the sharded descriptor database (one rank per GPU; ranks may be processes -- the
// torchrun layout bench.py uses -- or threads of one process, the "single process, 8 GPUs" layout of a ROS node).
#include "comm.cuh"

#include <dlfcn.h>

#include <mutex>

namespace cb {

const NcclApi* nccl_api() {
  static NcclApi api;
  static int state = 0;  // 0 = not tried, 1 = ok, -1 = failed
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (state == 1) return &api;
  if (state == -1) {
    fail(CB_ENODEVICE, "NCCL is not available (libnccl.so.2 could not be loaded)");
    return nullptr;
  }
  void* h = nullptr;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    state = -1;
    fail(CB_ENODEVICE, "NCCL is not available: %s", dlerror());
    return nullptr;
  }
  bool ok = true;
#define CB_SYM(field, sym)                                     \
  *reinterpret_cast<void**>(&api.field) = dlsym(h, sym);       \
  ok = ok && api.field != nullptr
  CB_SYM(GetUniqueId, "ncclGetUniqueId");
  CB_SYM(CommInitRank, "ncclCommInitRank");
  CB_SYM(CommDestroy, "ncclCommDestroy");
  CB_SYM(AllGather, "ncclAllGather");
  CB_SYM(GetErrorString, "ncclGetErrorString");
  CB_SYM(GetVersion, "ncclGetVersion");
#undef CB_SYM
  if (!ok) {
    state = -1;
    fail(CB_ENODEVICE, "libnccl is missing one of the entry points this library binds");
    return nullptr;
  }
  state = 1;
  return &api;
}

}  // namespace cb

extern "C" {

int cb_comm_get_unique_id(uint8_t* id_out) {
  if (!id_out) return cb::fail(CB_EINVAL, "id_out is NULL");
  static_assert(sizeof(ncclUniqueId) == CB_COMM_ID_BYTES, "CB_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
  const cb::NcclApi* api = cb::nccl_api();
  if (!api) return CB_ENODEVICE;
  ncclUniqueId id;
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return cb::fail(CB_ECUDA, "ncclGetUniqueId: %s", api->GetErrorString(r));
  memcpy(id_out, &id, sizeof(id));
  return CB_OK;
}

int cb_comm_create(cb_comm** out, const uint8_t* id, int rank, int world, int device) {
  if (!out) return cb::fail(CB_EINVAL, "out is NULL");
  *out = nullptr;
  if (!id || world < 1 || rank < 0 || rank >= world) return cb::fail(CB_EINVAL, "bad communicator arguments (rank %d of %d)", rank, world);
  int rc = cb::select_device(device, nullptr);
  if (rc) return rc;
  const cb::NcclApi* api = cb::nccl_api();
  if (!api) return CB_ENODEVICE;
  cb::DeviceGuard g(device);
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  cb_comm* c = new cb_comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
  ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
  if (r != ncclSuccess) {
    delete c;
    return cb::fail(CB_ECUDA, "ncclCommInitRank(rank %d of %d): %s", rank, world, api->GetErrorString(r));
  }
  *out = c;
  return CB_OK;
}

int cb_comm_destroy(cb_comm* c) {
  if (!c) return CB_OK;
  const cb::NcclApi* api = cb::nccl_api();
  if (api && c->comm) {
    cb::DeviceGuard g(c->device);
    api->CommDestroy(c->comm);
  }
  delete c;
  return CB_OK;
}

int cb_comm_rank(const cb_comm* c) { return c ? c->rank : -1; }
int cb_comm_world(const cb_comm* c) { return c ? c->world : -1; }

int cb_comm_nccl_version(void) {
  const cb::NcclApi* api = cb::nccl_api();
  if (!api) return -1;
  int v = 0;
  return api->GetVersion(&v) == ncclSuccess ? v : -1;
}

}  // extern "C"
