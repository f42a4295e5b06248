// comm.cu -- NCCL plumbing of the row-partitioned engine: one communicator per process / GPU, collectives
// enqueued on the engine's CUDA stream.  The reference is single-GPU (no NCCL/MPI anywhere); this is
// new work behind the same C ABI style as the kernels.
//
// NCCL is bound at RUN time (dlopen of libnccl.so.2) so that libgcn_b200.so keeps loading on machines without NCCL
// and, inside a PyTorch process, re-uses the NCCL library torch already mapped instead of dragging in a second one.
#include <dlfcn.h>
#include <nccl.h>  // types and enums only; no symbol of libnccl is linked

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

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

constexpr int kMaxWorld = 16;

// peer-memory slab exchange (NVLink / NVSwitch): every rank owns two gather buffers and a flag word per peer, opened
// in every other process through CUDA IPC
struct P2PPeers {
  float *gather[2][kMaxWorld];
  uint32_t *flags[kMaxWorld];
  int world, rank;
};

struct HaloOffsets {
  int64_t off[kMaxWorld + 1];  // rows for peer p: d_halo_rows[off[p] .. off[p + 1])
};

struct gcnb_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // slab gather state
  float *gather_local[2] = {nullptr, nullptr};
  uint32_t *flags_local = nullptr;
  unsigned int *d_ticket = nullptr;
  int64_t gather_floats = 0;
  bool p2p = false;
  P2PPeers peers{};
  void *opened[3 * kMaxWorld] = {nullptr};
  int n_opened = 0;
  uint32_t epoch = 0;
  // halo exchange (gcnb_comm_halo_setup): the rows of this rank's slab that each peer's row block references
  bool halo = false;
  int64_t halo_block = 0, halo_rows_local = 0, halo_send_total = 0, halo_need_total = 0;
  uint32_t *d_halo_rows = nullptr;  // [halo_send_total] local row ids, grouped by destination rank
  HaloOffsets halo_off{};
};

namespace {

// rank r's slab -> slot r of EVERY rank's gather buffer (its own included), then one flag per peer: "slab `epoch` of
// rank r has landed".  The last CTA to finish publishes the flags (system-scope fence before, release stores).
__global__ void __launch_bounds__(256)
p2p_push_kernel(const float4 *__restrict__ slab, int64_t n4, P2PPeers peers, int buf, int64_t slot_off4, uint32_t epoch,
                unsigned int *__restrict__ ticket) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(slab + i);
    for (int p = 0; p < peers.world; p++) reinterpret_cast<float4 *>(peers.gather[buf][p])[slot_off4 + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < peers.world) {
      uint32_t *f = peers.flags[threadIdx.x] + peers.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    }
    if (threadIdx.x == 0) *ticket = 0;
  }
}

// Halo exchange: the same buffers, flags and slot layout as p2p_push_kernel (row i of rank r lives at row r * block + i of
// every rank's buffer, so column ids need no translation and nothing downstream changes) -- but a peer receives only the
// rows its row block references; the rest of its copy of this rank's slot is never written and never read.  The rank's
// own slot gets the whole slab (local copy).  One thread per float4 of a row; consecutive list entries are mostly
// consecutive rows, so the NVLink stores coalesce.
__global__ void __launch_bounds__(256)
p2p_push_rows_kernel(const float4 *__restrict__ slab, int64_t n4_self, int d4, const uint32_t *__restrict__ rows, HaloOffsets off,
                     P2PPeers peers, int buf, int64_t slot_off4, uint32_t epoch, unsigned int *__restrict__ ticket) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float4 *self = reinterpret_cast<float4 *>(peers.gather[buf][peers.rank]) + slot_off4;
  for (int64_t i = t0; i < n4_self; i += stride) self[i] = __ldg(slab + i);
  const int64_t total = off.off[peers.world] * d4;
  int p = 0;
  for (int64_t t = t0; t < total; t += stride) {
    const int64_t e = t / d4;
    const int c = (int)(t - e * d4);
    while (e >= off.off[p + 1]) p++;  // entries are grouped by destination; t only grows
    const int64_t k = (int64_t)__ldg(rows + e) * d4 + c;
    reinterpret_cast<float4 *>(peers.gather[buf][p])[slot_off4 + k] = __ldg(slab + k);
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < peers.world) {
      uint32_t *f = peers.flags[threadIdx.x] + peers.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    }
    if (threadIdx.x == 0) *ticket = 0;
  }
}

// bit c of the mask <- 1 for every column id c of the rank's CSR block
__global__ void halo_mask_kernel(const uint32_t *__restrict__ indices, int64_t nnz, uint32_t *__restrict__ mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t c = __ldg(indices + i);
    const uint32_t bit = 1u << (c & 31);
    if (!(mask[c >> 5] & bit)) atomicOr(mask + (c >> 5), bit);
  }
}

// copy-engine variant of the push: the slabs travel as peer cudaMemcpyAsync calls (no SM, no LSU traffic -- the SMs are
// busy with the staged windows of the rank's own slab meanwhile); this kernel, next in stream order, publishes the flags
__global__ void p2p_flag_kernel(P2PPeers peers, uint32_t epoch) {
  __threadfence_system();
  if (threadIdx.x < peers.world) {
    uint32_t *f = peers.flags[threadIdx.x] + peers.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
}

// wait until every rank's slab `epoch` has landed in this rank's buffer
__global__ void p2p_wait_kernel(const uint32_t *__restrict__ flags, int world, uint32_t epoch) {
  if (threadIdx.x < world) {
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
    } while ((int32_t)(v - epoch) < 0);
  }
  __threadfence_system();
}

}  // namespace

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
    int rc = nccl_rc(a.CommInitRank(&c->comm, world, id, rank));
    if (rc) {
      delete c;
      return rc;
    }
    // NCCL sets its channels up lazily at the first collective (hundreds of milliseconds): pay for it here, where the
    // communicator is created, not inside the first model's constructor
    // ... and peer access to the other GPUs of the node (the slab exchange maps every peer's buffers with CUDA IPC;
    // enabling access lazily inside cudaIpcOpenMemHandle costs tens of milliseconds per peer: 0.3 s at 8 ranks)
    {
      int cur = 0, count = 0;
      if (cudaGetDevice(&cur) == cudaSuccess && cudaGetDeviceCount(&count) == cudaSuccess)
        for (int d = 0; d < count; d++) {
          int can = 0;
          if (d != cur && cudaDeviceCanAccessPeer(&can, cur, d) == cudaSuccess && can) cudaDeviceEnablePeerAccess(d, 0);
        }
      cudaGetLastError();  // already enabled / not supported: not an error here
    }
    uint32_t *d_one = nullptr;
    if (cudaMalloc((void **)&d_one, 4 * (size_t)world) == cudaSuccess) {
      cudaMemset(d_one, 0, 4 * (size_t)world);
      rc = nccl_rc(a.AllReduce(d_one, d_one, 1, ncclUint32, ncclSum, c->comm, nullptr));
      if (!rc) rc = nccl_rc(a.AllGather(d_one + rank, d_one, 1, ncclUint32, c->comm, nullptr));
      cudaStreamSynchronize(nullptr);
      cudaFree(d_one);
      if (rc) {
        nccl().CommDestroy(c->comm);
        delete c;
        return rc;
      }
    }
  }
  *out = c;
  return 0;
}

int gcnb_comm_destroy(gcnb_comm *c) {
  if (!c) return 0;
  cudaDeviceSynchronize();
  for (int i = 0; i < c->n_opened; i++) cudaIpcCloseMemHandle(c->opened[i]);
  cudaFree(c->gather_local[0]);
  cudaFree(c->gather_local[1]);
  cudaFree(c->flags_local);
  cudaFree(c->d_ticket);
  cudaFree(c->d_halo_rows);
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
  return 0;
}

// Allocates the slab-gather buffers (two, `gather_floats` each) and tries to open every peer's buffers through CUDA
// IPC.  If any rank cannot (no peer access, IPC unavailable), ALL ranks fall back to ncclAllGather into the local
// buffer -- still a device collective, decided collectively so that no rank waits for a push that never comes.
int gcnb_comm_gather_setup(gcnb_comm *c, int64_t gather_floats) {
  if (!c || gather_floats <= 0) return GCNB_E_BADARG;
  if (c->gather_local[0] && c->gather_floats >= gather_floats) return 0;
  if (c->gather_local[0]) {
    // a later, larger model on the same communicator: start over (every rank takes this branch -- sizes are equal on
    // all ranks -- and the handle exchange below is a barrier; nobody pushes for the finished model any more)
    GCNB_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < c->n_opened; i++) cudaIpcCloseMemHandle(c->opened[i]);
    c->n_opened = 0;
    cudaFree(c->gather_local[0]);
    cudaFree(c->gather_local[1]);
    cudaFree(c->flags_local);
    cudaFree(c->d_ticket);
    c->gather_local[0] = c->gather_local[1] = nullptr;
    c->flags_local = nullptr;
    c->d_ticket = nullptr;
    c->p2p = false;
    c->epoch = 0;
    cudaFree(c->d_halo_rows);
    c->d_halo_rows = nullptr;
    c->halo = false;
  }
  c->gather_floats = gather_floats;
  for (int b = 0; b < 2; b++) {
    GCNB_CHECK(cudaMalloc((void **)&c->gather_local[b], (size_t)gather_floats * 4));
    GCNB_CHECK(cudaMemset(c->gather_local[b], 0, (size_t)gather_floats * 4));
  }
  GCNB_CHECK(cudaMalloc((void **)&c->flags_local, kMaxWorld * 4));
  GCNB_CHECK(cudaMemset(c->flags_local, 0, kMaxWorld * 4));
  GCNB_CHECK(cudaMalloc((void **)&c->d_ticket, 4));
  GCNB_CHECK(cudaMemset(c->d_ticket, 0, 4));
  c->peers.world = c->world;
  c->peers.rank = c->rank;
  if (c->world == 1) return 0;
  const char *env = getenv("GCNB_P2P");  // GCNB_P2P=0: force the NCCL all-gather (A/B measurements)
  int ok = (c->world <= kMaxWorld) && !(env && atoi(env) == 0);
  // exchange the IPC handles (3 per rank) with one all-gather of raw bytes
  struct Handles {
    cudaIpcMemHandle_t h[3];
  };
  static_assert(sizeof(Handles) % 4 == 0, "handles travel as float words");
  Handles mine{};
  if (ok) {
    ok = cudaIpcGetMemHandle(&mine.h[0], c->gather_local[0]) == cudaSuccess &&
         cudaIpcGetMemHandle(&mine.h[1], c->gather_local[1]) == cudaSuccess &&
         cudaIpcGetMemHandle(&mine.h[2], c->flags_local) == cudaSuccess;
    if (!ok) cudaGetLastError();
  }
  Handles *d_all = nullptr;
  GCNB_CHECK(cudaMalloc((void **)&d_all, sizeof(Handles) * c->world));
  GCNB_CHECK(cudaMemcpy(d_all + c->rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice));
  int rc = nccl_rc(nccl().AllGather(d_all + c->rank, d_all, sizeof(Handles) / 4, ncclFloat, c->comm, nullptr));
  std::vector<Handles> all((size_t)c->world);
  if (!rc) rc = (int)cudaMemcpy(all.data(), d_all, sizeof(Handles) * c->world, cudaMemcpyDeviceToHost);
  cudaFree(d_all);
  if (rc) return rc;
  for (int p = 0; p < c->world && ok; p++) {
    if (p == c->rank) {
      c->peers.gather[0][p] = c->gather_local[0];
      c->peers.gather[1][p] = c->gather_local[1];
      c->peers.flags[p] = c->flags_local;
      continue;
    }
    void *ptr[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < 3 && ok; k++) {
      ok = cudaIpcOpenMemHandle(&ptr[k], all[p].h[k], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      if (ok) c->opened[c->n_opened++] = ptr[k];
      else cudaGetLastError();
    }
    c->peers.gather[0][p] = (float *)ptr[0];
    c->peers.gather[1][p] = (float *)ptr[1];
    c->peers.flags[p] = (uint32_t *)ptr[2];
  }
  // collective decision: everybody pushes or nobody does
  uint32_t *d_ok = nullptr;
  uint32_t h_ok = ok ? 1u : 0u;
  GCNB_CHECK(cudaMalloc((void **)&d_ok, 4));
  GCNB_CHECK(cudaMemcpy(d_ok, &h_ok, 4, cudaMemcpyHostToDevice));
  rc = nccl_rc(nccl().AllReduce(d_ok, d_ok, 1, ncclUint32, ncclSum, c->comm, nullptr));
  if (!rc) rc = (int)cudaMemcpy(&h_ok, d_ok, 4, cudaMemcpyDeviceToHost);
  cudaFree(d_ok);
  if (rc) return rc;
  c->p2p = h_ok == (uint32_t)c->world;
  return 0;
}

int gcnb_comm_gather_mode(const gcnb_comm *c) { return c ? (c->world == 1 ? 0 : (c->p2p ? 2 : 1)) : 0; }

// *d_full_out = [world x count_per_rank] floats: slab of rank r at r * count_per_rank.  Peer-memory mode: this rank's
// slab is stored straight into every rank's buffer over NVLink and a flag per (writer, reader) pair tells the reader
// when it has landed -- one copy kernel + one wait kernel, no ring / tree schedule.  Two buffers alternate: a rank can
// be at most one exchange ahead of a peer (it needed that peer's previous slab to get there), so the buffer it
// overwrites has been consumed.
int gcnb_comm_gather_slabs_f32(gcnb_comm *c, const float *d_slab, int64_t count_per_rank, const float **d_full_out,
                               gcnb_stream_t s) {
  return gcnb_comm_gather_slabs_ex_f32(c, d_slab, count_per_rank, d_full_out, 0, s);
}

int gcnb_comm_gather_slabs_ex_f32(gcnb_comm *c, const float *d_slab, int64_t count_per_rank, const float **d_full_out,
                                  int overlapped, gcnb_stream_t s) {
  if (!c || !d_slab || !d_full_out || count_per_rank <= 0 || !c->gather_local[0]) return GCNB_E_BADARG;
  if (count_per_rank * c->world > c->gather_floats || count_per_rank % 4 || ((uintptr_t)d_slab % 16)) return GCNB_E_BADARG;
  cudaStream_t st = as_stream(s);
  const uint32_t epoch = ++c->epoch;
  const int buf = (int)(epoch & 1u);
  float *full = c->gather_local[buf];
  *d_full_out = full;
  if (c->world == 1) {
    GCNB_CHECK(cudaMemcpyAsync(full, d_slab, (size_t)count_per_rank * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  if (!c->p2p) return nccl_rc(nccl().AllGather(d_slab, full, (size_t)count_per_rank, ncclFloat, c->comm, st));
  // exchange overlapped with compute on another stream and large enough to amortise 1 + world API calls: copy engines
  static const int dma_env = [] {
    const char *e = getenv("GCNB_P2P_DMA");  // tuning probe: 0 never, 1 always, unset = overlapped exchanges >= 4 MB
    return e ? atoi(e) : -1;
  }();
  if (c->halo && c->halo_block > 0 && count_per_rank % c->halo_block == 0 && (count_per_rank / c->halo_block) % 4 == 0) {
    const int d4 = (int)(count_per_rank / c->halo_block / 4);
    const int64_t n4_self = c->halo_rows_local * d4, work = std::max(n4_self, c->halo_send_total * d4);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)std::max(1, gcnb::device_info().sm_count) * 2));
    p2p_push_rows_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(d_slab), n4_self, d4, c->d_halo_rows, c->halo_off,
                                                 c->peers, buf, (int64_t)c->rank * (count_per_rank / 4), epoch, c->d_ticket);
    GCNB_LAUNCH_CHECK();
    p2p_wait_kernel<<<1, 32, 0, st>>>(c->flags_local, c->world, epoch);
    GCNB_LAUNCH_CHECK();
    return 0;
  }
  const bool dma = dma_env == 1 || (dma_env < 0 && overlapped && count_per_rank * 4 >= (4 << 20));
  if (dma) {
    for (int k = 0; k < c->world; k++) {
      const int p = (c->rank + 1 + k) % c->world;  // start with the next rank: no two ranks target one peer at once
      GCNB_CHECK(cudaMemcpyAsync(c->peers.gather[buf][p] + (size_t)c->rank * count_per_rank, d_slab,
                                 (size_t)count_per_rank * 4, cudaMemcpyDeviceToDevice, st));
    }
    p2p_flag_kernel<<<1, 32, 0, st>>>(c->peers, epoch);
    GCNB_LAUNCH_CHECK();
    p2p_wait_kernel<<<1, 32, 0, st>>>(c->flags_local, c->world, epoch);
    GCNB_LAUNCH_CHECK();
    return 0;
  }
  const int64_t n4 = count_per_rank / 4;
  const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)std::max(1, gcnb::device_info().sm_count) * 2);
  p2p_push_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(d_slab), n4, c->peers, buf,
                                          (int64_t)c->rank * n4, epoch, c->d_ticket);
  GCNB_LAUNCH_CHECK();
  p2p_wait_kernel<<<1, 32, 0, st>>>(c->flags_local, c->world, epoch);
  GCNB_LAUNCH_CHECK();
  return 0;
}

// Which rows of rank `rank`'s block does each peer reference?  masks = [world][words] bit masks over the slot-layout column
// space (bit c of masks[p]: rank p's row block holds column c).  Pure host logic (tests/test_halo_cpu.py).
int gcnb_halo_lists_from_masks(const uint32_t *masks, int world, int rank, int64_t words, int64_t block, int64_t rows_local,
                               uint32_t **rows_out, int64_t *off_out /* [world + 1] */) {
  if (!masks || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || block <= 0 || rows_local < 0 ||
      rows_local > block || words * 32 < (int64_t)world * block || !rows_out || !off_out)
    return GCNB_E_BADARG;
  std::vector<uint32_t> rows;
  const int64_t base = (int64_t)rank * block;
  for (int p = 0; p < world; p++) {
    off_out[p] = (int64_t)rows.size();
    if (p == rank) continue;
    const uint32_t *m = masks + (size_t)p * words;
    for (int64_t i = 0; i < rows_local; i++) {
      const int64_t c = base + i;
      if (m[c >> 5] >> (c & 31) & 1u) rows.push_back((uint32_t)i);
    }
  }
  off_out[world] = (int64_t)rows.size();
  uint32_t *out = (uint32_t *)malloc(std::max<size_t>(4, rows.size() * 4));
  if (!out) return GCNB_E_UNSUPPORTED;
  if (!rows.empty()) memcpy(out, rows.data(), rows.size() * 4);
  *rows_out = out;
  return 0;
}

// Halo exchange for the slab gather (graphs whose [N x d] matrix is too large to ship whole).  From the
// column ids of this rank's CSR block (slot layout: column c = row c % block of rank c / block) every rank marks the
// columns it references, the marks are all-gathered (n_global / 8 bytes per rank), and each rank keeps, per peer, the list
// of its OWN rows that peer references.  From then on gcnb_comm_gather_slabs*_f32 calls whose slab is [block x dim] push
// only those rows.  Sender-side decision: the lists are used when they hold at most max_fraction of what the full push
// sends (GCNB_HALO=0 / 1 forces it off / on); a receiver gets every row it references either way.  Peer-memory mode only
// (the NCCL all-gather fallback ships whole slabs).  info: {active, rows sent per exchange, rows of the full push
// ((world - 1) x rows_local), rows this rank references in other ranks' blocks}.
int gcnb_comm_halo_setup(gcnb_comm *c, const uint32_t *d_indices, int64_t nnz, int64_t rows_local, int64_t block,
                         double max_fraction, int64_t info[4], gcnb_stream_t s) {
  if (!c || (nnz > 0 && !d_indices) || nnz < 0 || block <= 0 || rows_local < 0 || rows_local > block) return GCNB_E_BADARG;
  if (info) info[0] = info[1] = info[2] = info[3] = 0;
  cudaFree(c->d_halo_rows);
  c->d_halo_rows = nullptr;
  c->halo = false;
  if (c->world == 1 || c->world > kMaxWorld) return 0;
  cudaStream_t st = as_stream(s);
  const int64_t words = ((int64_t)c->world * block + 31) / 32;
  uint32_t *d_masks = nullptr;
  GCNB_CHECK(cudaMalloc((void **)&d_masks, (size_t)words * 4 * c->world));
  uint32_t *mine = d_masks + (size_t)c->rank * words;
  GCNB_CHECK(cudaMemsetAsync(mine, 0, (size_t)words * 4, st));
  if (nnz > 0) {
    halo_mask_kernel<<<(int)std::min<int64_t>((nnz + 255) / 256, 4096), 256, 0, st>>>(d_indices, nnz, mine);
    if (cudaPeekAtLastError() != cudaSuccess) {
      cudaFree(d_masks);
      return (int)cudaGetLastError();
    }
  }
  int rc = nccl_rc(nccl().AllGather(mine, d_masks, (size_t)words, ncclUint32, c->comm, st));
  std::vector<uint32_t> masks((size_t)words * c->world);
  if (!rc) rc = (int)cudaMemcpyAsync(masks.data(), d_masks, masks.size() * 4, cudaMemcpyDeviceToHost, st);
  if (!rc) rc = (int)cudaStreamSynchronize(st);
  cudaFree(d_masks);
  if (rc) return rc;
  uint32_t *rows = nullptr;
  HaloOffsets off{};
  rc = gcnb_halo_lists_from_masks(masks.data(), c->world, c->rank, words, block, rows_local, &rows, off.off);
  if (rc) return rc;
  int64_t need = 0;  // rows of OTHER ranks' blocks that this rank references
  {
    const uint32_t *m = masks.data() + (size_t)c->rank * words;
    const int64_t lo = (int64_t)c->rank * block, hi = lo + block;
    for (int64_t w = 0; w < words; w++) {
      uint32_t v = m[w];
      while (v) {
        const int64_t col = w * 32 + __builtin_ctz(v);
        v &= v - 1;
        if (col < lo || col >= hi) need++;
      }
    }
  }
  const int64_t total = off.off[c->world], full = (int64_t)(c->world - 1) * rows_local;
  const char *env = getenv("GCNB_HALO");
  const bool want = env ? atoi(env) != 0 : (double)total <= max_fraction * (double)full;
  if (info) {
    info[1] = total;
    info[2] = full;
    info[3] = need;
  }
  if (want && c->p2p) {
    rc = (int)cudaMalloc((void **)&c->d_halo_rows, std::max<size_t>(4, (size_t)total * 4));
    if (!rc && total > 0) rc = (int)cudaMemcpy(c->d_halo_rows, rows, (size_t)total * 4, cudaMemcpyHostToDevice);
    if (!rc) {
      c->halo = true;
      c->halo_block = block;
      c->halo_rows_local = rows_local;
      c->halo_send_total = total;
      c->halo_need_total = need;
      c->halo_off = off;
      if (info) info[0] = 1;
    }
  }
  free(rows);
  return rc;
}

int gcnb_comm_halo_active(const gcnb_comm *c) { return c && c->halo ? 1 : 0; }

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
