// dist.cu -- NCCL plumbing of the mesh-partitioned evaluation behind the C ABI (ad4sm_dist_*, include/ad4sm_b200.h).
//
// The reference has no multi-device path (its only parallelism on this path is Threads.@threads over elements,
// /root/reference/src/elasticelements.jl:216); north_star asks for "interface-DOF contributions exchanged with NCCL over
// NVLink, ϕ reduced".  NCCL is bound at run time (dlopen of libnccl.so.2) so that the library has no link-time
// dependency: inside a PyTorch process the already-loaded torch NCCL is used, in a Julia / C host the system one.
#include <dlfcn.h>
#include <nccl.h>

#include <string>

#include "internal.h"

namespace ad4sm {

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};
NcclApi g_nccl;

bool nccl_load() {
  if (g_nccl.handle) return true;
  if (!g_nccl.error.empty()) return false;
  const char* names[] = {getenv("AD4SM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    g_nccl.error = "libnccl.so.2 not found (set AD4SM_NCCL_LIB): multi-GPU evaluation needs NCCL";
    return false;
  }
#define SYM(field, name)                                              \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
  if (!g_nccl.field) {                                                \
    g_nccl.error = std::string("NCCL symbol missing: ") + name;       \
    return false;                                                     \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(AllGather, "ncclAllGather")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.handle = h;
  return true;
}
std::string nccl_err(const char* what, ncclResult_t r) { return std::string(what) + ": " + g_nccl.GetErrorString(r); }
}  // namespace

struct DistComm {
  ncclComm_t comm;
  int nranks, rank, device;
};

int dist_unique_id(void* out, std::string* err) {
  static_assert(sizeof(ncclUniqueId) == 128, "AD4SM_NCCL_ID_BYTES");
  if (!nccl_load()) {
    *err = g_nccl.error;
    return -1;
  }
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) {
    *err = nccl_err("ncclGetUniqueId", r);
    return -1;
  }
  memcpy(out, &id, sizeof(id));
  return 0;
}

DistComm* dist_comm_init(int nranks, int rank, const void* id_bytes, int device, std::string* err) {
  if (!nccl_load()) {
    *err = g_nccl.error;
    return nullptr;
  }
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t c;
  ncclResult_t r = g_nccl.CommInitRank(&c, nranks, id, rank);
  if (r != ncclSuccess) {
    *err = nccl_err("ncclCommInitRank", r);
    return nullptr;
  }
  return new DistComm{c, nranks, rank, device};
}
void dist_comm_destroy(DistComm* c) {
  if (!c) return;
  if (g_nccl.handle) g_nccl.CommDestroy(c->comm);
  delete c;
}
int dist_comm_size(const DistComm* c) { return c->nranks; }
int dist_comm_rank(const DistComm* c) { return c->rank; }

// one grouped neighbour exchange: every send and receive of every peer and array in ONE ncclGroup (no ordering deadlock)
int dist_exchange(DistComm* c, int n, const DistXfer* x, cudaStream_t s, std::string* err) {
  ncclResult_t r = g_nccl.GroupStart();
  for (int i = 0; i < n && r == ncclSuccess; i++) {
    if (x[i].count == 0) continue;
    r = x[i].send ? g_nccl.Send(x[i].buf, x[i].count, ncclDouble, x[i].peer, c->comm, s)
                  : g_nccl.Recv(x[i].buf, x[i].count, ncclDouble, x[i].peer, c->comm, s);
  }
  const ncclResult_t r2 = g_nccl.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) {
    *err = nccl_err("ncclSend/ncclRecv", r);
    return -1;
  }
  return 0;
}
int dist_allgather(DistComm* c, const double* src, double* dst, size_t count, cudaStream_t s, std::string* err) {
  const ncclResult_t r = g_nccl.AllGather(src, dst, count, ncclDouble, c->comm, s);
  if (r != ncclSuccess) {
    *err = nccl_err("ncclAllGather", r);
    return -1;
  }
  return 0;
}

// dst[k][:] = src[rows[k]][:]  (rows of `width` doubles; width even -> 16-byte copies)
__global__ void __launch_bounds__(256) gather_rows_kernel(const double* src, const long long* rows, long long n_rows, int width,
                                                          double* dst) {
  const int w2 = width >> 1;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rows * w2) return;
  const long long k = t / w2;
  const int c = (int)(t - k * w2);
  reinterpret_cast<double2*>(dst + k * width)[c] = reinterpret_cast<const double2*>(src + rows[k] * width)[c];
}
__global__ void __launch_bounds__(256) gather_rows_kernel1(const double* src, const long long* rows, long long n_rows, int width,
                                                           double* dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rows * width) return;
  const long long k = t / width;
  const int c = (int)(t - k * width);
  dst[k * width + c] = src[rows[k] * width + c];
}
cudaError_t launch_gather_rows(const double* src, const long long* rows, long long n_rows, int width, double* dst, cudaStream_t s) {
  if (n_rows <= 0) return cudaSuccess;
  if (width % 2 == 0) {
    const long long n = n_rows * (width / 2);
    gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, rows, n_rows, width, dst);
  } else {
    const long long n = n_rows * width;
    gather_rows_kernel1<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, rows, n_rows, width, dst);
  }
  return cudaGetLastError();
}
// ϕ = Σ_g ϕ_g in rank order (fixed association: run-to-run reproducible)
__global__ void phi_rank_sum_kernel(const double* allphi, int n, double* out) {
  double s = allphi[0];
  for (int g = 1; g < n; g++) s += allphi[g];
  *out = s;
}
cudaError_t launch_phi_rank_sum(const double* allphi, int n, double* out, cudaStream_t s) {
  phi_rank_sum_kernel<<<1, 1, 0, s>>>(allphi, n, out);
  return cudaGetLastError();
}

}  // namespace ad4sm
