// comm.cu -- NCCL plumbing of the row-partitioned engine: one communicator per process / GPU, collectives
// enqueued on the engine's CUDA stream.  The reference is single-GPU (no NCCL/MPI anywhere); this is
// new work behind the same C ABI style as the kernels.
//
// NCCL is bound at RUN time (dlopen of libnccl.so.2) so that libgcn_b200.so keeps loading on machines without NCCL
// and, inside a PyTorch process, re-uses the NCCL library torch already mapped instead of dragging in a second one.
#include <dlfcn.h>
#include <nccl.h>  // types and enums only; no symbol of libnccl is linked

#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    auto sym = [&](const char *n) { return dlsym(api.handle, n); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce &&
             api.GroupStart && api.GroupEnd;
  });
  return api;
}

int nccl_rc(ncclResult_t r) {
  if (r == ncclSuccess) return 0;
  const NcclApi &a = nccl();
  fprintf(stderr, "libgcn_b200: NCCL error %d: %s\n", (int)r, a.GetErrorString ? a.GetErrorString(r) : "?");
  return GCNB_E_COMM;
}

}  // namespace

struct gcnb_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

extern "C" {

int gcnb_comm_unique_id(void *out_id) {
  if (!out_id) return GCNB_E_BADARG;
  NcclApi &a = nccl();
  if (!a.ok) return GCNB_E_COMM;
  ncclUniqueId id;
  const int rc = nccl_rc(a.GetUniqueId(&id));
  if (rc) return rc;
  static_assert(sizeof(id) == GCNB_COMM_ID_BYTES, "ncclUniqueId size");
  memcpy(out_id, &id, sizeof(id));
  return 0;
}

int gcnb_comm_create(int rank, int world, const void *id_bytes, gcnb_comm **out) {
  if (!out || world < 1 || rank < 0 || rank >= world) return GCNB_E_BADARG;
  auto *c = new gcnb_comm();
  c->rank = rank;
  c->world = world;
  if (world > 1) {
    if (!id_bytes) {
      delete c;
      return GCNB_E_BADARG;
    }
    NcclApi &a = nccl();
    if (!a.ok) {
      delete c;
      return GCNB_E_COMM;
    }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    const int rc = nccl_rc(a.CommInitRank(&c->comm, world, id, rank));
    if (rc) {
      delete c;
      return rc;
    }
  }
  *out = c;
  return 0;
}

int gcnb_comm_destroy(gcnb_comm *c) {
  if (!c) return 0;
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
  return 0;
}

int gcnb_comm_rank(const gcnb_comm *c) { return c ? c->rank : 0; }
int gcnb_comm_world(const gcnb_comm *c) { return c ? c->world : 1; }

int gcnb_comm_all_gather_f32(gcnb_comm *c, const float *d_send, float *d_recv, int64_t count_per_rank, gcnb_stream_t s) {
  if (!c || !d_send || !d_recv || count_per_rank < 0) return GCNB_E_BADARG;
  if (c->world == 1) {
    if (d_send != d_recv)
      GCNB_CHECK(cudaMemcpyAsync(d_recv, d_send, (size_t)count_per_rank * 4, cudaMemcpyDeviceToDevice, as_stream(s)));
    return 0;
  }
  return nccl_rc(nccl().AllGather(d_send, d_recv, (size_t)count_per_rank, ncclFloat, c->comm, as_stream(s)));
}

int gcnb_comm_all_reduce_sum(gcnb_comm *c, void *d_buf, int64_t count, int is_u32, gcnb_stream_t s) {
  if (!c || !d_buf || count < 0) return GCNB_E_BADARG;
  if (c->world == 1 || count == 0) return 0;
  return nccl_rc(nccl().AllReduce(d_buf, d_buf, (size_t)count, is_u32 ? ncclUint32 : ncclFloat, ncclSum, c->comm, as_stream(s)));
}

int gcnb_comm_group_start(gcnb_comm *c) { return (c && c->world > 1) ? nccl_rc(nccl().GroupStart()) : 0; }
int gcnb_comm_group_end(gcnb_comm *c) { return (c && c->world > 1) ? nccl_rc(nccl().GroupEnd()) : 0; }

}  // extern "C"
