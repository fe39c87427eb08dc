// Multi-GPU feature exchange of the frame-sharded extract+match path: the ONE collective of the
// path, behind the C-ABI ("multi-GPU feature exchange" section of pgb200.h).
//
// Frames shard over the GPUs of a box in contiguous blocks; extraction needs no communication; matching frame t needs
// the keypoints + descriptors of frame t-1, so exactly one fixed-size per-frame record crosses each block boundary.
// pgb_allgather_feats is an NCCL all-gather over NVLink / NVSwitch of whatever block of records every rank contributes:
// the boundary record alone (62 KB per rank: what the matcher needs; bench.py, optical_trajectories) or a rank's whole
// block (the full feature table on every rank, for callers that want it).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded by a host process such as PyTorch wins,
// else the system library), so libpgb200.so itself has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace {

// The slice of nccl.h this file needs (NCCL 2.x ABI: stable since 2.0).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // ncclSuccess = 0
typedef int ncclDataType_t;  // ncclInt8 = 0 (ncclChar), ncclUint8 = 1
constexpr ncclDataType_t kNcclUint8 = 1;

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  void* handle = nullptr;
  std::string error;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + s; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return &api;
}

int nccl_fail(const char* what, ncclResult_t r) {
  NcclApi* a = nccl();
  return pgb::fail(PGB_ERR_CUDA, "%s: NCCL error %d (%s)", what, (int)r, a->GetErrorString ? a->GetErrorString(r) : "?");
}

}  // namespace

namespace {

// One frame's features as a contiguous record: [0,16) count + padding | [16, 16 + 28*cap) keypoints | descriptors (32*cap)
// at the next 16-byte boundary.  Pack / unpack move between the record and the per-frame arrays pgb_orb_extract writes.
__host__ __device__ inline size_t record_desc_off(int cap) { return (16 + (size_t)cap * 28 + 15) & ~(size_t)15; }

__global__ void k_record_copy(int cap, int toRecord, uint32_t* __restrict__ kps, uint32_t* __restrict__ desc, int* __restrict__ count,
                              uint32_t* __restrict__ rec) {
  const int nK = cap * 7, nD = cap * 8;
  uint32_t* rk = rec + 4;
  uint32_t* rd = rec + record_desc_off(cap) / 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nK + nD; i += gridDim.x * blockDim.x) {
    uint32_t* a = i < nK ? kps + i : desc + (i - nK);
    uint32_t* b = i < nK ? rk + i : rd + (i - nK);
    if (toRecord) *b = *a; else *a = *b;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (toRecord) { rec[0] = (uint32_t)*count; rec[1] = rec[2] = rec[3] = 0u; } else *count = (int)rec[0];
  }
}

}  // namespace

struct pgb_comm {
  int device = 0, rank = 0, size = 1;
  ncclComm_t comm = nullptr;
};

using namespace pgb;

extern "C" {

int pgb_comm_unique_id(uint8_t id[PGB_COMM_ID_BYTES]) {
  if (!id) return fail(PGB_ERR_INVALID, "null id");
  NcclApi* a = nccl();
  if (!a->error.empty()) return fail(PGB_ERR_CUDA, "%s", a->error.c_str());
  ncclUniqueId u;
  ncclResult_t r = a->GetUniqueId(&u);
  if (r) return nccl_fail("ncclGetUniqueId", r);
  static_assert(sizeof u == PGB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  memcpy(id, &u, sizeof u);
  return PGB_OK;
}

pgb_comm* pgb_comm_create(int device, int rank, int n_ranks, const uint8_t id[PGB_COMM_ID_BYTES]) {
  if (!id || n_ranks <= 0 || rank < 0 || rank >= n_ranks) { fail(PGB_ERR_INVALID, "pgb_comm_create: invalid argument"); return nullptr; }
  NcclApi* a = nccl();
  if (!a->error.empty()) { fail(PGB_ERR_CUDA, "%s", a->error.c_str()); return nullptr; }
  if (use_device(device)) return nullptr;
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  pgb_comm* c = new pgb_comm;
  c->device = device; c->rank = rank; c->size = n_ranks;
  ncclResult_t r = a->CommInitRank(&c->comm, n_ranks, u, rank);
  if (r) { nccl_fail("ncclCommInitRank", r); delete c; return nullptr; }
  return c;
}

int pgb_comm_create_all(int n, const int* devices, pgb_comm** out) {
  if (n <= 0 || !devices || !out) return fail(PGB_ERR_INVALID, "pgb_comm_create_all: invalid argument");
  NcclApi* a = nccl();
  if (!a->error.empty()) return fail(PGB_ERR_CUDA, "%s", a->error.c_str());
  for (int i = 0; i < n; i++) {
    int rc = use_device(devices[i]);
    if (rc) return rc;
  }
  std::vector<ncclComm_t> comms(n);
  ncclResult_t r = a->CommInitAll(comms.data(), n, devices);
  if (r) return nccl_fail("ncclCommInitAll", r);
  for (int i = 0; i < n; i++) {
    out[i] = new pgb_comm;
    out[i]->device = devices[i]; out[i]->rank = i; out[i]->size = n; out[i]->comm = comms[i];
  }
  return PGB_OK;
}

void pgb_comm_destroy(pgb_comm* c) {
  if (!c) return;
  if (c->comm) {
    cudaSetDevice(c->device);
    nccl()->CommDestroy(c->comm);
  }
  delete c;
}

int pgb_comm_rank(const pgb_comm* c) { return c ? c->rank : PGB_ERR_INVALID; }
int pgb_comm_size(const pgb_comm* c) { return c ? c->size : PGB_ERR_INVALID; }

int pgb_comm_nccl_version(void) {
  NcclApi* a = nccl();
  int v = 0;
  if (!a->error.empty() || !a->GetVersion || a->GetVersion(&v)) return 0;
  return v;
}

size_t pgb_frame_record_bytes(int cap) { return cap > 0 ? record_desc_off(cap) + (size_t)cap * 32 : 0; }

// the launch must happen with the buffers' device current (host threads that drive several GPUs call these too)
static int select_device_of(const void* p) {
  cudaPointerAttributes a;
  PGB_CUDA(cudaPointerGetAttributes(&a, p));
  if (a.type != cudaMemoryTypeDevice) return fail(PGB_ERR_INVALID, "frame record buffers must be device memory");
  PGB_CUDA(cudaSetDevice(a.device));
  return PGB_OK;
}

int pgb_frame_record_pack(const pgb_keypoint* kps, const uint8_t* desc, const int32_t* counts, int frame, int cap, void* record,
                          void* stream) {
  if (!kps || !desc || !counts || !record || frame < 0 || cap <= 0) return fail(PGB_ERR_INVALID, "pgb_frame_record_pack: invalid argument");
  if (int rc = select_device_of(record)) return rc;
  k_record_copy<<<8, 256, 0, (cudaStream_t)stream>>>(cap, 1, (uint32_t*)(kps + (size_t)frame * cap), (uint32_t*)(desc + (size_t)frame * cap * 32),
                                                      (int*)(counts + frame), (uint32_t*)record);
  PGB_CHECK_LAUNCH();
  return PGB_OK;
}

int pgb_frame_record_unpack(const void* record, pgb_keypoint* kps, uint8_t* desc, int32_t* counts, int frame, int cap, void* stream) {
  if (!kps || !desc || !counts || !record || frame < 0 || cap <= 0) return fail(PGB_ERR_INVALID, "pgb_frame_record_unpack: invalid argument");
  if (int rc = select_device_of(record)) return rc;
  k_record_copy<<<8, 256, 0, (cudaStream_t)stream>>>(cap, 0, (uint32_t*)(kps + (size_t)frame * cap), (uint32_t*)(desc + (size_t)frame * cap * 32),
                                                      (int*)(counts + frame), (uint32_t*)const_cast<void*>(record));
  PGB_CHECK_LAUNCH();
  return PGB_OK;
}

int pgb_allgather_feats(pgb_comm* c, const void* send, void* recv, size_t bytes_per_rank, void* stream) {
  if (!c || !send || !recv) return fail(PGB_ERR_INVALID, "pgb_allgather_feats: null argument");
  if (bytes_per_rank == 0) return PGB_OK;
  NvtxRange range("pgb_allgather_feats");
  PGB_CUDA(cudaSetDevice(c->device));
  ncclResult_t r = nccl()->AllGather(send, recv, bytes_per_rank, kNcclUint8, c->comm, (cudaStream_t)stream);
  if (r) return nccl_fail("ncclAllGather", r);
  return PGB_OK;
}

}  // extern "C"
