// NCCL halo sum (see halo.hpp).  NCCL is loaded with dlopen("libnccl.so.2"): inside a
// torch.distributed process this resolves to the library torch already loaded.
#include "halo.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

namespace mrhyde_b200 {

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string why;
};

NcclApi& api() {
  static NcclApi A;
  if (A.handle || !A.why.empty()) return A;
  A.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!A.handle) { A.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return A; }
  auto sym = [&](const char* n) { void* p = dlsym(A.handle, n); if (!p && A.why.empty()) A.why = std::string("libnccl.so.2 lacks ") + n; return p; };
  A.GetUniqueId = (decltype(A.GetUniqueId))sym("ncclGetUniqueId");
  A.CommInitRank = (decltype(A.CommInitRank))sym("ncclCommInitRank");
  A.CommDestroy = (decltype(A.CommDestroy))sym("ncclCommDestroy");
  A.Send = (decltype(A.Send))sym("ncclSend");
  A.Recv = (decltype(A.Recv))sym("ncclRecv");
  A.AllGather = (decltype(A.AllGather))sym("ncclAllGather");
  A.GroupStart = (decltype(A.GroupStart))sym("ncclGroupStart");
  A.GroupEnd = (decltype(A.GroupEnd))sym("ncclGroupEnd");
  A.GetErrorString = (decltype(A.GetErrorString))sym("ncclGetErrorString");
  A.ok = A.why.empty();
  return A;
}

#define NCCL_TRY(call)                                                                        \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess) { err = std::string(#call) + ": " + api().GetErrorString(r_); return false; } \
  } while (0)
#define CU_TRY(call)                                                                          \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } \
  } while (0)

__global__ void pack_kernel(const double* __restrict__ src, const int64_t* __restrict__ pos, int64_t n, double* __restrict__ dst) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[pos[i]];
}
__global__ void unpack_add_kernel(const double* __restrict__ src, const int64_t* __restrict__ pos, int64_t n, double* __restrict__ dst) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && pos[i] >= 0) dst[pos[i]] += src[i];
}
// residual and matrix values of one source rank in one launch: [0, n_res) -> res, [n_res, n_res + n_jac) -> jac
__global__ void unpack_add2_kernel(const double* __restrict__ src, const int64_t* __restrict__ pos_res, int64_t n_res, double* __restrict__ res,
                                   const int64_t* __restrict__ pos_jac, int64_t n_jac, double* __restrict__ jac) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_res) { if (pos_res[i] >= 0) res[pos_res[i]] += src[i]; }
  else if (i < n_res + n_jac) { const int64_t p = pos_jac[i - n_res]; if (p >= 0) jac[p] += src[i]; }
}

// ---- p2p transport -------------------------------------------------------------------------------------------------
// Region of rank r (one cudaMalloc, mapped by its neighbours): [arrive[nranks] | ack[nranks]] 64-bit flags, then for every source
// rank two receive slabs (double-buffered by the parity of the call counter), each laid out [res | pad to even | jac].
struct P2PPeerDev {
  // what goes to this peer
  const int64_t* send_res_pos; const int64_t* send_jac_pos;
  int64_t n_send_res, n_send_jac, send_res_first, send_jac_first, send_jac_off;
  double* remote_slab[2];
  unsigned long long* remote_arrive;      // peer's arrive[me]
  unsigned long long* local_ack;          // my ack[peer]: the peer is done reading what I stored in call `value`
  // what comes from this peer
  const int64_t* recv_res_pos; const int64_t* recv_jac_pos;
  int64_t n_recv_res, n_recv_jac, recv_jac_off;
  const double* local_slab[2];
  unsigned long long* local_arrive;       // my arrive[peer]
  unsigned long long* remote_ack;         // peer's ack[me]
  unsigned long long* remote_shift;       // peer's shift[me]: doubles between the slab's matrix part and its first value (0 here; the in-kernel push may use 1)
  const unsigned long long* local_shift;  // my shift[peer]
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// One launch per halo sum, at most one CTA per SM (all CTAs are resident, so the flag waits cannot starve a CTA that has not started).
//   phase 1  every CTA stores its share of the ghost rows into the owners' slabs; the last CTA to finish a peer raises arrive[me] there
//   phase 2  source rank by source rank (ascending: fixed summation order): wait for arrive[source], add the slab, last CTA acknowledges
// counters[0 .. n): CTAs that finished the push to peer i; counters[n .. 2n): CTAs that finished the add from peer i (both reset by
// their last CTA); counters[2n]: running ticket of the grid barrier between two sources.
__global__ void __launch_bounds__(1024) halo_p2p_kernel(const P2PPeerDev* __restrict__ peers, int n, unsigned long long epoch, double* __restrict__ res,
                                                        double* __restrict__ jac, unsigned* counters, int pushed_peer) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int par = (int)(epoch & 1ull);
  for (int i = 0; i < n; ++i) {
    const P2PPeerDev& P = peers[i];
    const int64_t nr = res ? P.n_send_res : 0, nj = jac ? P.n_send_jac : 0;
    if (nr + nj == 0 || i == pushed_peer) continue;   // pushed_peer: the assembly kernel already stored these rows and raised the flag
    if (threadIdx.x == 0) while (ld_acquire_sys(P.local_ack) + 2ull < epoch) {}   // the slab of this parity was read two calls ago
    __syncthreads();
    double* slab = P.remote_slab[par];
    if (P.send_res_first >= 0) for (int64_t k = tid; k < nr; k += nth) slab[k] = res[P.send_res_first + k];
    else for (int64_t k = tid; k < nr; k += nth) slab[k] = res[P.send_res_pos[k]];
    double* sj = slab + P.send_jac_off;
    if (P.send_jac_first >= 0) {
      const double* src = jac + P.send_jac_first;
#pragma unroll 4
      for (int64_t k = tid; k < nj; k += nth) sj[k] = __ldcs(src + k);   // independent loads in flight, posted stores over NVLink
    } else {
#pragma unroll 4
      for (int64_t k = tid; k < nj; k += nth) sj[k] = jac[P.send_jac_pos[k]];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(&counters[i], 1u) == gridDim.x - 1) {
        counters[i] = 0u;
        __threadfence_system();
        st_release_sys(P.remote_shift, 0ull);
        st_release_sys(P.remote_arrive, epoch);
      }
    }
  }
  bool first = true;
  for (int i = 0; i < n; ++i) {
    const P2PPeerDev& P = peers[i];
    const int64_t nr = res ? P.n_recv_res : 0, nj = jac ? P.n_recv_jac : 0;
    if (nr + nj == 0) continue;
    if (threadIdx.x == 0) {
      if (!first) {   // rows shared with several ranks: the previous source must be added everywhere before this one starts
        const unsigned ticket = atomicAdd(&counters[2 * n], 1u);
        const unsigned target = (ticket / gridDim.x + 1u) * gridDim.x;
        while (*(volatile unsigned*)&counters[2 * n] < target) {}
        __threadfence();
      }
      while (ld_acquire_sys(P.local_arrive) < epoch) {}
    }
    first = false;
    __syncthreads();
    const double* slab = P.local_slab[par];
    for (int64_t k = tid; k < nr; k += nth) { const int64_t q = P.recv_res_pos[k]; if (q >= 0) res[q] += __ldcg(slab + k); }
    const double* sj = slab + P.recv_jac_off + (int64_t)ld_acquire_sys(P.local_shift);
#pragma unroll 4
    for (int64_t k = tid; k < nj; k += nth) { const int64_t q = __ldcs(P.recv_jac_pos + k); if (q >= 0) jac[q] += __ldcg(sj + k); }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(&counters[n + i], 1u) == gridDim.x - 1) {
        counters[n + i] = 0u;
        st_release_sys(P.remote_ack, epoch);
      }
    }
  }
}

template <class T>
bool to_device(T** d, const std::vector<T>& h, std::string& err) {
  CU_TRY(cudaMalloc((void**)d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) CU_TRY(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return true;
}

}  // namespace

bool halo_unique_id(uint8_t* id128, std::string& err) {
  NcclApi& A = api();
  if (!A.ok) { err = A.why; return false; }
  ncclUniqueId id;
  NCCL_TRY(A.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(id128, &id, 128);
  return true;
}

HaloExchange::~HaloExchange() {
  for (auto& p : peers_) {
    cudaFree(p.d_send_res); cudaFree(p.d_send_jac); cudaFree(p.d_recv_res); cudaFree(p.d_recv_jac);
    cudaFree(p.d_sendbuf); cudaFree(p.d_recvbuf);
  }
  for (size_t p = 0; p < p2p_remote_.size(); ++p) if (p2p_remote_[p]) cudaIpcCloseMemHandle(p2p_remote_[p]);
  cudaFree(p2p_region_); cudaFree(p2p_peers_dev_); cudaFree(p2p_counters_); cudaFree(p2p_push_counter_);
  if (side_) { cudaStreamSynchronize(side_); cudaStreamDestroy(side_); cudaEventDestroy(ev_ready_); cudaEventDestroy(ev_done_); }
  if (comm_ && api().ok) api().CommDestroy((ncclComm_t)comm_);
}

bool HaloExchange::init(const uint8_t* id128, int rank, int nranks, std::string& err) {
  NcclApi& A = api();
  if (!A.ok) { err = A.why; return false; }
  if (rank < 0 || rank >= nranks) { err = "halo: rank out of range"; return false; }
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t c;
  NCCL_TRY(A.CommInitRank(&c, nranks, id, rank));
  comm_ = c; rank_ = rank; nranks_ = nranks;
  return true;
}

// Variable-size all-to-all of int64 payloads through grouped send/recv (host vectors in, host vectors out).
static bool exchange_lists(NcclApi& A, ncclComm_t comm, int rank, int nranks, const std::vector<std::vector<int64_t>>& out_lists,
                           std::vector<std::vector<int64_t>>& in_lists, std::string& err) {
  // counts
  std::vector<int64_t> cnt((size_t)nranks), allcnt((size_t)nranks * nranks);
  for (int p = 0; p < nranks; ++p) cnt[(size_t)p] = (int64_t)out_lists[(size_t)p].size();
  int64_t *d_cnt = nullptr, *d_all = nullptr;
  CU_TRY(cudaMalloc(&d_cnt, nranks * sizeof(int64_t)));
  CU_TRY(cudaMalloc(&d_all, (size_t)nranks * nranks * sizeof(int64_t)));
  CU_TRY(cudaMemcpy(d_cnt, cnt.data(), nranks * sizeof(int64_t), cudaMemcpyHostToDevice));
  NCCL_TRY(A.AllGather(d_cnt, d_all, (size_t)nranks, ncclInt64, comm, 0));
  CU_TRY(cudaStreamSynchronize(0));
  CU_TRY(cudaMemcpy(allcnt.data(), d_all, allcnt.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
  cudaFree(d_cnt); cudaFree(d_all);
  in_lists.assign((size_t)nranks, {});
  std::vector<int64_t*> dsend((size_t)nranks, nullptr), drecv((size_t)nranks, nullptr);
  for (int p = 0; p < nranks; ++p) {
    if (p == rank) continue;
    const int64_t ns = cnt[(size_t)p], nr = allcnt[(size_t)p * nranks + rank];
    in_lists[(size_t)p].resize((size_t)nr);
    if (ns) { CU_TRY(cudaMalloc(&dsend[(size_t)p], ns * sizeof(int64_t))); CU_TRY(cudaMemcpy(dsend[(size_t)p], out_lists[(size_t)p].data(), ns * sizeof(int64_t), cudaMemcpyHostToDevice)); }
    if (nr) CU_TRY(cudaMalloc(&drecv[(size_t)p], nr * sizeof(int64_t)));
  }
  NCCL_TRY(A.GroupStart());
  for (int p = 0; p < nranks; ++p) {
    if (p == rank) continue;
    if (!out_lists[(size_t)p].empty()) NCCL_TRY(A.Send(dsend[(size_t)p], out_lists[(size_t)p].size(), ncclInt64, p, comm, 0));
    if (!in_lists[(size_t)p].empty()) NCCL_TRY(A.Recv(drecv[(size_t)p], in_lists[(size_t)p].size(), ncclInt64, p, comm, 0));
  }
  NCCL_TRY(A.GroupEnd());
  CU_TRY(cudaStreamSynchronize(0));
  for (int p = 0; p < nranks; ++p) {
    if (!in_lists[(size_t)p].empty()) CU_TRY(cudaMemcpy(in_lists[(size_t)p].data(), drecv[(size_t)p], in_lists[(size_t)p].size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
    cudaFree(dsend[(size_t)p]); cudaFree(drecv[(size_t)p]);
  }
  return true;
}

bool HaloExchange::setup(int64_t n_rows, int64_t n_owned, int64_t n_cols, const int64_t* row_gids, const int64_t* rowptr, const int32_t* colind, std::string& err) {
  // row_gids doubles as the column gid table: [0,n_rows) rows, [n_rows,n_cols) column-only ghosts
  NcclApi& A = api();
  ncclComm_t comm = (ncclComm_t)comm_;
  ready_ = false;
  // 1. every rank learns who owns which global row: all-to-all of the owned gid lists
  std::vector<std::vector<int64_t>> out((size_t)nranks_), in;
  for (int p = 0; p < nranks_; ++p) if (p != rank_) out[(size_t)p].assign(row_gids, row_gids + n_owned);
  if (!exchange_lists(A, comm, rank_, nranks_, out, in, err)) return false;
  std::unordered_map<int64_t, int> owner;
  owner.reserve((size_t)n_rows * 2);
  for (int p = 0; p < nranks_; ++p) for (int64_t g : in[(size_t)p]) owner[g] = p;
  // 2. for each peer: my ghost rows it owns -> message [row gid, ncols, col gids...]; value order = same walk
  std::vector<std::vector<int64_t>> desc((size_t)nranks_);
  std::vector<std::vector<int64_t>> send_res((size_t)nranks_), send_jac((size_t)nranks_);
  for (int64_t r = n_owned; r < n_rows; ++r) {
    auto it = owner.find(row_gids[r]);
    if (it == owner.end()) { err = "halo: ghost row " + std::to_string(row_gids[r]) + " is owned by no rank"; return false; }
    const int p = it->second;
    desc[(size_t)p].push_back(row_gids[r]);
    desc[(size_t)p].push_back(rowptr[r + 1] - rowptr[r]);
    send_res[(size_t)p].push_back(r);
    for (int64_t q = rowptr[r]; q < rowptr[r + 1]; ++q) { desc[(size_t)p].push_back(row_gids[colind[q]]); send_jac[(size_t)p].push_back(q); }
  }
  std::vector<std::vector<int64_t>> rdesc;
  if (!exchange_lists(A, comm, rank_, nranks_, desc, rdesc, err)) return false;
  // 3. destination positions on the owner
  std::unordered_map<int64_t, int64_t> my_row;  // gid -> local column id (rows, then column-only ghosts)
  my_row.reserve((size_t)n_cols * 2);
  for (int64_t r = 0; r < n_cols; ++r) my_row[row_gids[r]] = r;
  peers_.assign((size_t)nranks_, Peer());
  for (int p = 0; p < nranks_; ++p) {
    if (p == rank_) continue;
    std::vector<int64_t> recv_res, recv_jac;
    const auto& d = rdesc[(size_t)p];
    size_t k = 0;
    while (k < d.size()) {
      const int64_t gid = d[k++], nc = d[k++];
      auto it = my_row.find(gid);
      if (it == my_row.end() || it->second >= n_owned) { err = "halo: received a row this rank does not own"; return false; }
      const int64_t r = it->second;
      recv_res.push_back(r);
      for (int64_t c = 0; c < nc; ++c) {
        const int64_t cg = d[k++];
        int64_t pos = -1;
        auto ic = my_row.find(cg);
        if (ic != my_row.end()) {
          const int32_t lc = (int32_t)ic->second;
          const int32_t* b = colind + rowptr[r];
          const int32_t* e = colind + rowptr[r + 1];
          const int32_t* f = std::lower_bound(b, e, lc);
          if (f != e && *f == lc) pos = rowptr[r] + (f - b);
        }
        if (pos < 0) { err = "halo: owner row lacks a column present in the ghost copy (graphs are inconsistent)"; return false; }
        recv_jac.push_back(pos);
      }
    }
    Peer& P = peers_[(size_t)p];
    P.n_send_res = (int64_t)send_res[(size_t)p].size(); P.n_send_jac = (int64_t)send_jac[(size_t)p].size();
    P.n_recv_res = (int64_t)recv_res.size(); P.n_recv_jac = (int64_t)recv_jac.size();
    // ghost rows come last in the overlapped numbering, so what goes to a neighbour is usually one contiguous slice of res and
    // one of the value array: those are sent in place, without a pack kernel
    auto contiguous = [](const std::vector<int64_t>& v) {
      for (size_t i = 1; i < v.size(); ++i) if (v[i] != v[i - 1] + 1) return false;
      return !v.empty();
    };
    P.send_res_first = contiguous(send_res[(size_t)p]) ? send_res[(size_t)p][0] : -1;
    P.send_jac_first = contiguous(send_jac[(size_t)p]) ? send_jac[(size_t)p][0] : -1;
    if (!to_device(&P.d_send_res, send_res[(size_t)p], err) || !to_device(&P.d_send_jac, send_jac[(size_t)p], err) ||
        !to_device(&P.d_recv_res, recv_res, err) || !to_device(&P.d_recv_jac, recv_jac, err)) return false;
    CU_TRY(cudaMalloc(&P.d_sendbuf, std::max<int64_t>(1, P.n_send_res + P.n_send_jac) * sizeof(double)));
    CU_TRY(cudaMalloc(&P.d_recvbuf, std::max<int64_t>(1, P.n_recv_res + P.n_recv_jac) * sizeof(double)));
  }
  if (!setup_p2p(err)) return false;
  ready_ = true;
  return true;
}

// Collective.  Every rank allocates its region, publishes the IPC handle and the slab sizes it expects, maps the regions of the
// ranks it sends to, and the ranks agree (all-gather of a success word) whether the p2p transport is used.
bool HaloExchange::setup_p2p(std::string& err) {
  NcclApi& A = api();
  ncclComm_t comm = (ncclComm_t)comm_;
  p2p_ = false;
  const char* env = getenv("MRHYDE_B200_HALO_TRANSPORT");
  const std::string want = env && *env ? std::string(env) : transport_;
  if (want != "auto" && want != "p2p" && want != "nccl") { err = "halo transport must be auto|p2p|nccl"; return false; }
  // layout of my region
  const size_t flag_bytes = (size_t)nranks_ * 3 * sizeof(unsigned long long);   // arrive[], ack[], shift[]
  std::vector<int64_t> slab_off((size_t)nranks_, -1), slab_len((size_t)nranks_, 0);
  size_t total = (flag_bytes + 255) & ~(size_t)255;
  for (int p = 0; p < nranks_; ++p) {
    if (p == rank_) continue;
    const Peer& P = peers_[(size_t)p];
    if (P.n_recv_res + P.n_recv_jac == 0) continue;
    slab_len[(size_t)p] = ((P.n_recv_res + 1) & ~(int64_t)1) + P.n_recv_jac + 2;   // + room for the in-kernel push's one-double shift
    slab_off[(size_t)p] = (int64_t)total;
    total += (((size_t)slab_len[(size_t)p] * sizeof(double) + 255) & ~(size_t)255) * 2;
  }
  int ok = want == "nccl" ? 0 : 1;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok) {
    if (cudaMalloc(&p2p_region_, total) != cudaSuccess || cudaMemset(p2p_region_, 0, total) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, p2p_region_) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  }
  // publish: [ok, handle (64 bytes = 8 words), slab_off[nranks]]
  const size_t words = 1 + 8 + (size_t)nranks_;
  std::vector<int64_t> pub(words, 0), all(words * (size_t)nranks_, 0);
  pub[0] = ok;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(&pub[1], &mine, 64);
  for (int p = 0; p < nranks_; ++p) pub[9 + (size_t)p] = slab_off[(size_t)p];
  auto gather = [&](std::vector<int64_t>& in, std::vector<int64_t>& out) -> bool {
    int64_t *d_in = nullptr, *d_out = nullptr;
    CU_TRY(cudaMalloc(&d_in, in.size() * sizeof(int64_t)));
    CU_TRY(cudaMalloc(&d_out, out.size() * sizeof(int64_t)));
    CU_TRY(cudaMemcpy(d_in, in.data(), in.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    NCCL_TRY(A.AllGather(d_in, d_out, in.size(), ncclInt64, comm, 0));
    CU_TRY(cudaStreamSynchronize(0));
    CU_TRY(cudaMemcpy(out.data(), d_out, out.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
    cudaFree(d_in); cudaFree(d_out);
    return true;
  };
  if (!gather(pub, all)) return false;
  for (int p = 0; p < nranks_; ++p) if (!all[words * (size_t)p]) ok = 0;
  // map the regions of the ranks I exchange with
  p2p_remote_.assign((size_t)nranks_, nullptr);
  if (ok) {
    for (int p = 0; p < nranks_ && ok; ++p) {
      if (p == rank_) continue;
      const Peer& P = peers_[(size_t)p];
      if (P.n_send_res + P.n_send_jac + P.n_recv_res + P.n_recv_jac == 0) continue;
      cudaIpcMemHandle_t h;
      std::memcpy(&h, &all[words * (size_t)p + 1], 64);
      if (cudaIpcOpenMemHandle(&p2p_remote_[(size_t)p], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; p2p_remote_[(size_t)p] = nullptr; cudaGetLastError(); }
    }
  }
  std::vector<int64_t> okv(1, ok), okall((size_t)nranks_, 0);
  if (!gather(okv, okall)) return false;
  for (int p = 0; p < nranks_; ++p) if (!okall[(size_t)p]) ok = 0;
  if (!ok) {
    if (want == "p2p") { err = "halo transport p2p: a rank could not allocate, export or map a receive region (CUDA IPC / peer access)"; return false; }
    for (auto& r : p2p_remote_) if (r) { cudaIpcCloseMemHandle(r); r = nullptr; }
    cudaFree(p2p_region_); p2p_region_ = nullptr;
    return true;   // the nccl transport stays in place
  }
  // device-side peer table
  std::vector<P2PPeerDev> tab;
  p2p_tab_rank_.clear();
  p2p_slab_remote_.assign((size_t)nranks_, nullptr);
  p2p_slab_stride_.assign((size_t)nranks_, 0);
  for (int p = 0; p < nranks_; ++p) {   // ascending rank: the order the received values are added in
    if (p == rank_) continue;
    const Peer& P = peers_[(size_t)p];
    if (P.n_send_res + P.n_send_jac + P.n_recv_res + P.n_recv_jac == 0) continue;
    P2PPeerDev D;
    std::memset(&D, 0, sizeof(D));
    char* remote = (char*)p2p_remote_[(size_t)p];
    char* local = (char*)p2p_region_;
    D.send_res_pos = P.d_send_res; D.send_jac_pos = P.d_send_jac;
    D.n_send_res = P.n_send_res; D.n_send_jac = P.n_send_jac; D.send_res_first = P.send_res_first; D.send_jac_first = P.send_jac_first;
    D.send_jac_off = (P.n_send_res + 1) & ~(int64_t)1;
    const int64_t roff = all[words * (size_t)p + 9 + (size_t)rank_];   // where the peer keeps the slabs for what I send
    if (P.n_send_res + P.n_send_jac > 0) {
      if (roff < 0) { err = "halo p2p: the owner expects nothing from a rank that has ghost rows for it (inconsistent maps)"; return false; }
      const size_t len = (((size_t)(D.send_jac_off + P.n_send_jac + 2) * sizeof(double) + 255) & ~(size_t)255);
      D.remote_slab[0] = (double*)(remote + roff); D.remote_slab[1] = (double*)(remote + roff + len);
      p2p_slab_remote_[(size_t)p] = remote + roff; p2p_slab_stride_[(size_t)p] = len;
    }
    D.remote_arrive = (unsigned long long*)remote + rank_;
    D.local_ack = (unsigned long long*)local + nranks_ + p;
    D.recv_res_pos = P.d_recv_res; D.recv_jac_pos = P.d_recv_jac;
    D.n_recv_res = P.n_recv_res; D.n_recv_jac = P.n_recv_jac;
    D.recv_jac_off = (P.n_recv_res + 1) & ~(int64_t)1;
    if (slab_off[(size_t)p] >= 0) {
      const size_t len = (((size_t)slab_len[(size_t)p] * sizeof(double) + 255) & ~(size_t)255);
      D.local_slab[0] = (const double*)(local + slab_off[(size_t)p]); D.local_slab[1] = (const double*)(local + slab_off[(size_t)p] + len);
    }
    D.local_arrive = (unsigned long long*)local + p;
    D.remote_ack = (unsigned long long*)remote + nranks_ + rank_;
    D.remote_shift = (unsigned long long*)remote + 2 * nranks_ + rank_;
    D.local_shift = (const unsigned long long*)local + 2 * nranks_ + p;
    tab.push_back(D);
    p2p_tab_rank_.push_back(p);
  }
  p2p_active_ = (int)tab.size();
  CU_TRY(cudaMalloc(&p2p_peers_dev_, std::max<size_t>(1, tab.size()) * sizeof(P2PPeerDev)));
  if (!tab.empty()) CU_TRY(cudaMemcpy(p2p_peers_dev_, tab.data(), tab.size() * sizeof(P2PPeerDev), cudaMemcpyHostToDevice));
  CU_TRY(cudaMalloc((void**)&p2p_counters_, (2 * tab.size() + 1) * sizeof(unsigned)));
  CU_TRY(cudaMemset(p2p_counters_, 0, (2 * tab.size() + 1) * sizeof(unsigned)));
  CU_TRY(cudaMalloc((void**)&p2p_push_counter_, sizeof(unsigned)));
  CU_TRY(cudaMemset(p2p_push_counter_, 0, sizeof(unsigned)));
  int dev = 0, n_sm = 0;
  CU_TRY(cudaGetDevice(&dev));
  CU_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  int64_t most = 0;
  for (const P2PPeerDev& D : tab) most = std::max<int64_t>(most, std::max(D.n_send_res + D.n_send_jac, D.n_recv_res + D.n_recv_jac));
  p2p_grid_ = (int)std::max<int64_t>(1, std::min<int64_t>(n_sm, (most + 2047) / 2048));   // <= one CTA of 1024 threads per SM: all resident
  p2p_epoch_ = 0;
  CU_TRY(cudaDeviceSynchronize());
  p2p_ = true;
  return true;
}

bool HaloExchange::push_params(const double* res, const double* jac, int64_t n_owned, int64_t ghost_base, int n_push_chains, PushDev& X) {
  std::memset(&X, 0, sizeof(X));
  push_peer_ = -1;
  if (!p2p_ || !ready_ || !res || !jac || n_push_chains <= 0) return false;
  int owner = -1, n_owners = 0;
  for (int p = 0; p < nranks_; ++p) {
    if (p == rank_) continue;
    const Peer& P = peers_[(size_t)p];
    if (P.n_send_res + P.n_send_jac > 0) { owner = p; ++n_owners; }
  }
  if (n_owners != 1) return false;
  const Peer& P = peers_[(size_t)owner];
  if (P.send_res_first != n_owned || P.send_jac_first != ghost_base || !p2p_slab_remote_[(size_t)owner]) return false;
  int ti = -1;
  for (size_t k = 0; k < p2p_tab_rank_.size(); ++k) if (p2p_tab_rank_[k] == owner) ti = (int)k;
  if (ti < 0) return false;
  const unsigned long long epoch = p2p_epoch_ + 1;   // the call counter of the sum() that follows this assembly
  char* slab = p2p_slab_remote_[(size_t)owner] + (size_t)(epoch & 1ull) * p2p_slab_stride_[(size_t)owner];
  double* slab_jac = (double*)slab + ((P.n_send_res + 1) & ~(int64_t)1);
  // one double of shift gives the slab the 16-byte phase of the local array (bulk copies need matching phases)
  const int shift = (int)(((reinterpret_cast<uintptr_t>(jac) >> 3) + (uintptr_t)ghost_base - (reinterpret_cast<uintptr_t>(slab_jac) >> 3)) & 1u);
  char* remote = (char*)p2p_remote_[(size_t)owner];
  X.remote_jac = slab_jac + shift;
  X.remote_res = (double*)slab;
  X.res_base = res;
  X.remote_arrive = (unsigned long long*)remote + rank_;
  X.remote_shift = (unsigned long long*)remote + 2 * nranks_ + rank_;
  X.local_ack = (const unsigned long long*)p2p_region_ + nranks_ + owner;
  X.counter = p2p_push_counter_;
  X.epoch = epoch;
  X.ghost_base = ghost_base;
  X.n_owned = (int32_t)n_owned; X.n_push_chains = n_push_chains; X.shift = shift; X.enabled = 1;
  push_peer_ = ti; pushed_res_ = res; pushed_jac_ = jac;
  return true;
}

bool HaloExchange::sum_p2p(double* res, double* jac, cudaStream_t st, std::string& err) {
  ++p2p_epoch_;
  launches_ = 0;
  const int pushed = (push_peer_ >= 0 && res == pushed_res_ && jac == pushed_jac_) ? push_peer_ : -1;
  if (push_peer_ >= 0 && pushed < 0) { err = "halo: the assembly pushed its ghost rows for other arrays than the ones passed to halo_sum"; return false; }
  push_peer_ = -1;
  if (p2p_active_ == 0) return true;
  halo_p2p_kernel<<<p2p_grid_, 1024, 0, st>>>((const P2PPeerDev*)p2p_peers_dev_, p2p_active_, p2p_epoch_, res, jac, p2p_counters_, pushed);
  launches_ = 1;
  CU_TRY(cudaGetLastError());
  return true;
}

bool HaloExchange::can_start() const {
  if (!ready_) return false;
  for (int p = 0; p < nranks_; ++p) {
    const Peer& P = peers_[(size_t)p];
    if (p == rank_) continue;
    if ((P.n_send_res && P.send_res_first < 0) || (P.n_send_jac && P.send_jac_first < 0)) return false;   // would need a pack kernel
  }
  return true;
}

bool HaloExchange::start(double* res, double* jac, cudaStream_t st, std::string& err) {
  NcclApi& A = api();
  ncclComm_t comm = (ncclComm_t)comm_;
  if (!can_start() || !res || !jac) { err = "halo: start() needs contiguous ghost slices and both arrays"; return false; }
  if (!side_) {
    int prio_lo = 0, prio_hi = 0;
    CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    // highest priority: the few CTAs of the send/recv kernel should get onto the SMs as soon as CTAs of the assembly retire
    CU_TRY(cudaStreamCreateWithPriority(&side_, cudaStreamNonBlocking, prio_hi));
    CU_TRY(cudaEventCreateWithFlags(&ev_ready_, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&ev_done_, cudaEventDisableTiming));
  }
  CU_TRY(cudaEventRecord(ev_ready_, st));
  CU_TRY(cudaStreamWaitEvent(side_, ev_ready_, 0));
  NCCL_TRY(A.GroupStart());
  for (int p = 0; p < nranks_; ++p) {
    Peer& P = peers_[(size_t)p];
    if (p == rank_) continue;
    if (P.n_send_res) NCCL_TRY(A.Send(res + P.send_res_first, (size_t)P.n_send_res, ncclFloat64, p, comm, side_));
    if (P.n_send_jac) NCCL_TRY(A.Send(jac + P.send_jac_first, (size_t)P.n_send_jac, ncclFloat64, p, comm, side_));
    if (P.n_recv_res) NCCL_TRY(A.Recv(P.d_recvbuf, (size_t)P.n_recv_res, ncclFloat64, p, comm, side_));
    if (P.n_recv_jac) NCCL_TRY(A.Recv(P.d_recvbuf + P.n_recv_res, (size_t)P.n_recv_jac, ncclFloat64, p, comm, side_));
  }
  NCCL_TRY(A.GroupEnd());
  CU_TRY(cudaEventRecord(ev_done_, side_));
  started_ = true; started_res_ = res; started_jac_ = jac;
  return true;
}

bool HaloExchange::sum(double* res, double* jac, cudaStream_t st, std::string& err) {
  if (p2p_ && !started_) return sum_p2p(res, jac, st, err);
  NcclApi& A = api();
  ncclComm_t comm = (ncclComm_t)comm_;
  auto blocks = [](int64_t n) { return (unsigned)((n + 255) / 256); };
  int launched = 0;
  if (started_) {   // the exchange of these arrays is already in flight (start()): wait for it, then add on `st`
    if (res != started_res_ || jac != started_jac_) { err = "halo: sum() after start() must be given the same arrays"; return false; }
    started_ = false;
    CU_TRY(cudaStreamWaitEvent(st, ev_done_, 0));
    for (int p = 0; p < nranks_; ++p) {  // ascending source rank: fixed summation order
      Peer& P = peers_[(size_t)p];
      if (p == rank_) continue;
      if (P.n_recv_res + P.n_recv_jac) {
        unpack_add2_kernel<<<blocks(P.n_recv_res + P.n_recv_jac), 256, 0, st>>>(P.d_recvbuf, P.d_recv_res, P.n_recv_res, res, P.d_recv_jac, P.n_recv_jac, jac);
        ++launched;
      }
    }
    launches_ = launched;
    CU_TRY(cudaGetLastError());
    return true;
  }
  for (int p = 0; p < nranks_; ++p) {
    Peer& P = peers_[(size_t)p];
    if (p == rank_) continue;
    if (res && P.n_send_res && P.send_res_first < 0) pack_kernel<<<blocks(P.n_send_res), 256, 0, st>>>(res, P.d_send_res, P.n_send_res, P.d_sendbuf), ++launched;
    if (jac && P.n_send_jac && P.send_jac_first < 0) pack_kernel<<<blocks(P.n_send_jac), 256, 0, st>>>(jac, P.d_send_jac, P.n_send_jac, P.d_sendbuf + P.n_send_res), ++launched;
  }
  NCCL_TRY(A.GroupStart());
  for (int p = 0; p < nranks_; ++p) {
    Peer& P = peers_[(size_t)p];
    if (p == rank_) continue;
    // two messages per peer and direction (residual entries, matrix values); the receive buffer is laid out [res | jac]
    if (res && P.n_send_res) NCCL_TRY(A.Send(P.send_res_first >= 0 ? res + P.send_res_first : P.d_sendbuf, (size_t)P.n_send_res, ncclFloat64, p, comm, st));
    if (jac && P.n_send_jac) NCCL_TRY(A.Send(P.send_jac_first >= 0 ? jac + P.send_jac_first : P.d_sendbuf + P.n_send_res, (size_t)P.n_send_jac, ncclFloat64, p, comm, st));
    if (res && P.n_recv_res) NCCL_TRY(A.Recv(P.d_recvbuf, (size_t)P.n_recv_res, ncclFloat64, p, comm, st));
    if (jac && P.n_recv_jac) NCCL_TRY(A.Recv(P.d_recvbuf + P.n_recv_res, (size_t)P.n_recv_jac, ncclFloat64, p, comm, st));
  }
  NCCL_TRY(A.GroupEnd());
  for (int p = 0; p < nranks_; ++p) {  // ascending source rank: fixed summation order
    Peer& P = peers_[(size_t)p];
    if (p == rank_) continue;
    if (res && jac && P.n_recv_res + P.n_recv_jac) {
      unpack_add2_kernel<<<blocks(P.n_recv_res + P.n_recv_jac), 256, 0, st>>>(P.d_recvbuf, P.d_recv_res, P.n_recv_res, res, P.d_recv_jac, P.n_recv_jac, jac);
      ++launched;
      continue;
    }
    if (res && P.n_recv_res) unpack_add_kernel<<<blocks(P.n_recv_res), 256, 0, st>>>(P.d_recvbuf, P.d_recv_res, P.n_recv_res, res), ++launched;
    if (jac && P.n_recv_jac) unpack_add_kernel<<<blocks(P.n_recv_jac), 256, 0, st>>>(P.d_recvbuf + P.n_recv_res, P.d_recv_jac, P.n_recv_jac, jac), ++launched;
  }
  launches_ = launched;
  CU_TRY(cudaGetLastError());
  return true;
}

}  // namespace mrhyde_b200
