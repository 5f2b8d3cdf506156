// Multi-GPU halo sum: the replacement of Tpetra's Export(overlapped -> owned, ADD) for the residual
// vector and the Jacobian values (linearAlgebraInterface_matrix.hpp:233-237, _vector.hpp:56-66;
// maps built at linearAlgebraInterface_construct.hpp:159-207).
//
// Each rank assembles its own elements into its overlapped rows (owned rows first, then ghost rows).
// halo sum: the values of every ghost row are packed, sent to the owning rank with NCCL send/recv
// over NVLink, and added into the owner's row at positions matched through global column ids at
// set-up.  Contributions are added source rank by source rank in ascending order, so the result is
// reproducible.  NCCL is resolved with dlopen at run time (the library loads without it).
//
// Two transports (option "halo transport" = auto | p2p | nccl, or MRHYDE_B200_HALO_TRANSPORT):
//   p2p   one kernel per halo sum and no NCCL on the data path: every rank maps its neighbours' receive slabs (CUDA IPC over
//         NVLink), stores its ghost rows straight into the owner's slab, raises a flag there and then adds what its own neighbours
//         stored (halo_p2p_kernel).  Used when every rank of the communicator could map its peers (one node).
//   nccl  grouped ncclSend/ncclRecv + an add kernel per source (the round-1 path; also the fallback across nodes).
#pragma once
#include <cuda_runtime.h>

#include "kernel_abi.h"

#include <cstdint>
#include <string>
#include <vector>

namespace mrhyde_b200 {

bool halo_unique_id(uint8_t* id128, std::string& err);

class HaloExchange {
 public:
  HaloExchange() {}
  ~HaloExchange();
  bool init(const uint8_t* id128, int rank, int nranks, std::string& err);
  // collective: builds send/receive maps from the global row ids of every rank
  bool setup(int64_t n_rows, int64_t n_owned, int64_t n_cols, const int64_t* col_gids, const int64_t* rowptr, const int32_t* colind, std::string& err);
  void set_transport(const std::string& t) { transport_ = t; }   // before setup(): auto | p2p | nccl
  bool p2p() const { return p2p_; }
  // In-kernel push (kernel_abi.h: PushDev): when this rank's ghost rows go to ONE owner as the contiguous tails res[n_owned ..) and
  // jac[ghost_base ..), fills X for the assembly kernel that is about to write res / jac -- the next sum() on the same arrays then
  // skips its own copy phase.  Returns false (X.enabled = 0) when the layout or the transport does not allow it.
  bool push_params(const double* res, const double* jac, int64_t n_owned, int64_t ghost_base, int n_push_chains, PushDev& X);
  bool sum(double* res, double* jac, cudaStream_t st, std::string& err);
  // Overlapped form of sum(): start() is called once the ghost rows of res / jac are final on `st` (the caller may keep
  // launching work that does not touch them); the send/recv run on an internal stream.  sum() called afterwards with the same
  // arrays waits for them and adds the received values on `st`.  Ghost slices must be contiguous (in-place send) for start().
  bool can_start() const;
  bool start(double* res, double* jac, cudaStream_t st, std::string& err);
  // an exchange started earlier whose sum() was never called: `st` waits for it before the arrays are written again
  void drain(cudaStream_t st) { if (started_) { cudaStreamWaitEvent(st, ev_done_, 0); started_ = false; } }
  bool ready() const { return ready_; }
  int launches_per_sum() const { return launches_; }

 private:
  void* comm_ = nullptr;
  int rank_ = 0, nranks_ = 1;
  bool ready_ = false;
  int launches_ = 0;
  // overlapped exchange
  cudaStream_t side_ = nullptr;
  cudaEvent_t ev_ready_ = nullptr, ev_done_ = nullptr;
  double* started_res_ = nullptr; double* started_jac_ = nullptr;
  bool started_ = false;
  // per peer: what I send (positions in my res / jac arrays) and where received values are added
  struct Peer {
    int64_t n_send_res = 0, n_send_jac = 0, n_recv_res = 0, n_recv_jac = 0;
    int64_t send_res_first = -1, send_jac_first = -1;   // >= 0: the send positions are one contiguous slice starting here (sent in place)
    int64_t* d_send_res = nullptr; int64_t* d_send_jac = nullptr;   // source positions
    int64_t* d_recv_res = nullptr; int64_t* d_recv_jac = nullptr;   // destination positions (-1: column absent on the owner)
    double* d_sendbuf = nullptr; double* d_recvbuf = nullptr;
  };
  std::vector<Peer> peers_;
  // ---- p2p transport
  bool setup_p2p(std::string& err);
  bool sum_p2p(double* res, double* jac, cudaStream_t st, std::string& err);
  std::string transport_ = "auto";
  bool p2p_ = false;
  void* p2p_region_ = nullptr;            // my flags + receive slabs (cudaMalloc, exported with cudaIpcGetMemHandle)
  std::vector<void*> p2p_remote_;         // peers' regions mapped here (cudaIpcOpenMemHandle), nullptr where unused
  void* p2p_peers_dev_ = nullptr;         // P2PPeerDev[n_active]
  unsigned* p2p_counters_ = nullptr;      // CTA arrival counters of the kernel
  int p2p_active_ = 0, p2p_grid_ = 0;
  unsigned long long p2p_epoch_ = 0;
  unsigned* p2p_push_counter_ = nullptr;  // push chains of the assembly kernel that have finished
  int push_peer_ = -1;                    // index (device peer table) of the owner the last push_params prepared, -1: none pending
  const double* pushed_res_ = nullptr; const double* pushed_jac_ = nullptr;
  std::vector<int> p2p_tab_rank_;         // rank of every entry of the device peer table
  std::vector<char*> p2p_slab_remote_;    // per rank: the owner's slabs for what I send (parity 0; parity 1 follows at p2p_slab_stride_)
  std::vector<size_t> p2p_slab_stride_;
};

}  // namespace mrhyde_b200
