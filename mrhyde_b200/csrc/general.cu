// Device side of the general assembly path: kernel instantiation table, the pull (row-sum) kernel and the launch
// sequence of one assemble call (see general.hpp / general_kernel.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "general.hpp"
#include "general_dispatch.hpp"

namespace mrhyde_b200 {

void gen_device_part0(std::vector<GenDeviceKernels>& T);
void gen_device_part1(std::vector<GenDeviceKernels>& T);
void gen_device_part2(std::vector<GenDeviceKernels>& T);
void gen_device_part3(std::vector<GenDeviceKernels>& T);

namespace {

std::vector<GenDeviceKernels>& device_table() {
  static std::vector<GenDeviceKernels> T;
  if (T.empty()) { gen_device_part0(T); gen_device_part1(T); gen_device_part2(T); gen_device_part3(T); }
  return T;
}

// ---- pull: one group of G lanes per CSR row ---------------------------------------------------------------------
struct PullParams {
  const int32_t* row_order;
  const int64_t* contrib_ptr;
  const int32_t* contrib;       // scratch slot * N + local row of every contribution (ascending instance order)
  const int32_t* contrib_pos;   // row of the (de-duplicated) column-position table: pos_id[instance] * N + local row
  const uint16_t* pos;
  const double* elem_jac;
  const double* elem_res;
  GraphDev G;
  OutDev O;
  int64_t row_begin, row_end, n_owned;
  int32_t N, max_row_len;
  // mass pull: fixed rows are not skipped, instances >= mass_inst_end (boundary sides) are ignored, O.res receives the
  // diagonal vector (Jacobi diagonal, or the lumped row sums of |entries| when mass_mode == 2)
  int32_t mass_mode;
  int32_t lump;                 // Solver: lump mass -- every entry of an element row is added to the row's diagonal (assemblyManager_scatter.hpp:263-268)
  int64_t mass_inst_end;        // scratch slots >= this one hold boundary-side instances
};

// G lanes per row, NPL = ceil(N / G) columns per lane and instance; the loads of U instances are issued before the first
// is accumulated so that each lane has U * NPL element-matrix loads in flight (the kernel is latency-bound otherwise)
template <int G, int NPL>
__global__ void __launch_bounds__(256) gen_pull_kernel(const __grid_constant__ PullParams Q) {
  constexpr int U = 4;
  extern __shared__ __align__(16) double pull_smem[];
  const int tid = threadIdx.x, lane = tid % G, grp = tid / G;
  const int64_t k = Q.row_begin + (int64_t)blockIdx.x * (blockDim.x / G) + grp;
  const unsigned mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((tid % 32) / G * G));
  if (k >= Q.row_end) return;
  double* buf = pull_smem + (size_t)grp * Q.max_row_len;
  const int32_t r = __ldg(Q.row_order + k);
  const int64_t rs = __ldg(Q.G.rowptr + r);
  const int len = (int)(__ldg(Q.G.rowptr + r + 1) - rs);
  if (Q.mass_mode == 3) {   // applyMassMatrixFree: y(r) (+)= sum of the element vectors M_e x_e, volume instances, no fixed-dof handling
    if (lane == 0) {
      const int64_t c0 = __ldg(Q.contrib_ptr + k), c1 = __ldg(Q.contrib_ptr + k + 1);
      double s = 0.0;
      for (int64_t p = c0; p < c1; ++p) {
        const int64_t inst = __ldg(Q.contrib + p);
        if (inst / Q.N < Q.mass_inst_end) s += __ldcs(Q.elem_res + inst);   // inst = slot * N + row
      }
      Q.O.res[r] = (Q.O.accumulate ? Q.O.res[r] : 0.0) + s;
    }
    return;
  }
  if (Q.mass_mode) {
    const int N = Q.N;
    const int64_t c0 = __ldg(Q.contrib_ptr + k), c1 = __ldg(Q.contrib_ptr + k + 1);
    for (int t = lane; t < len; t += G) buf[t] = 0.0;
    __syncwarp(mask);
    double lumped = 0.0;
    for (int64_t p = c0; p < c1; ++p) {
      const int64_t inst = __ldg(Q.contrib + p);
      if (inst / N >= Q.mass_inst_end) continue;
      const int64_t ci = inst * N, pi = (int64_t)__ldg(Q.contrib_pos + p) * N;
      for (int c = lane; c < N; c += G) { const double v = __ldcs(Q.elem_jac + ci + c); buf[__ldg(Q.pos + pi + c)] += v; lumped += fabs(v); }
      __syncwarp(mask);
    }
    if (Q.O.jac) for (int t = lane; t < len; t += G) Q.O.jac[rs + t] = (Q.O.accumulate ? Q.O.jac[rs + t] : 0.0) + buf[t];
    if (Q.O.res) {
      for (int o = G / 2; o > 0; o >>= 1) lumped += __shfl_xor_sync(mask, lumped, o, G);
      double d = lumped;
      if (Q.mass_mode != 2) {
        d = 0.0;
        for (int t = 0; t < len; ++t) if (__ldg(Q.G.colind + rs + t) == r) d = buf[t];
      }
      if (lane == 0) Q.O.res[r] = (Q.O.accumulate ? Q.O.res[r] : 0.0) + d;
    }
    return;
  }
  if (__ldg(Q.G.fixed + r)) {   // strong-Dirichlet row: skipped by the scatter (scatter.hpp:208, 253); identity row when overwriting
    if (!Q.O.accumulate) {
      if (Q.O.res && lane == 0) Q.O.res[r] = 0.0;
      if (Q.O.jac)
        for (int t = lane; t < len; t += G) Q.O.jac[rs + t] = (__ldg(Q.G.colind + rs + t) == r && r < Q.n_owned && Q.O.diag_one) ? 1.0 : 0.0;
    }
    return;
  }
  const int N = Q.N;
  const int64_t c0 = __ldg(Q.contrib_ptr + k), c1 = __ldg(Q.contrib_ptr + k + 1);
  int dpos = -1;
  if (Q.lump) {   // position of the diagonal entry in this row
    for (int t = lane; t < len; t += G) if (__ldg(Q.G.colind + rs + t) == r) dpos = t;
    for (int o = G / 2; o > 0; o >>= 1) dpos = max(dpos, __shfl_xor_sync(mask, dpos, o, G));
  }
  if (Q.O.jac) {
    for (int t = lane; t < len; t += G) buf[t] = 0.0;
    __syncwarp(mask);
    for (int64_t p = c0; p < c1; p += U) {
      double v[U][NPL];
      int ps[U][NPL];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t ci = p + u < c1 ? (int64_t)__ldg(Q.contrib + p + u) * N : -1;
        const int64_t pi = p + u < c1 ? (int64_t)__ldg(Q.contrib_pos + p + u) * N : 0;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          const int c = lane + j * G;
          const bool ok = ci >= 0 && c < N;
          ps[u][j] = ok ? (int)__ldg(Q.pos + pi + c) : -1;
          v[u][j] = ok ? __ldcs(Q.elem_jac + ci + c) : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {   // ascending instance order, one instance at a time: distinct positions within an instance
        if (Q.lump) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NPL; ++j) if (ps[u][j] >= 0) s += v[u][j];
          for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(mask, s, o, G);
          if (lane == 0 && dpos >= 0) buf[dpos] += s;
        } else {
#pragma unroll
          for (int j = 0; j < NPL; ++j) if (ps[u][j] >= 0) buf[ps[u][j]] += v[u][j];
        }
        __syncwarp(mask);
      }
    }
    if (Q.O.accumulate) for (int t = lane; t < len; t += G) Q.O.jac[rs + t] += buf[t];
    else for (int t = lane; t < len; t += G) __stcs(Q.O.jac + rs + t, buf[t]);
  }
  if (Q.O.res && lane == 0) {
    double s = 0.0;
    for (int64_t p = c0; p < c1; ++p) s += __ldcs(Q.elem_res + __ldg(Q.contrib + p));
    Q.O.res[r] = (Q.O.accumulate ? Q.O.res[r] : 0.0) + (-s);
  }
}

template <class T>
struct Buf {
  T* p = nullptr;
  size_t n = 0;
  ~Buf() { if (p) cudaFree(p); }
  bool upload(const std::vector<T>& h, size_t* tot, std::string& err) {
    n = h.size();
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess && n) e = cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { err = std::string("general plan upload: ") + cudaGetErrorString(e); return false; }
    if (tot) *tot += bytes;
    return true;
  }
  bool alloc(size_t count, size_t* tot, std::string& err) {
    n = count;
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    const cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { err = std::string("general plan scratch: ") + cudaGetErrorString(e); return false; }
    if (tot) *tot += bytes;
    return true;
  }
};

struct SideDev {
  Buf<int32_t> items;
  Buf<double> geo_N, geo_dN, ref_tab, qwts;
};

}  // namespace

struct GeneralPlanDev {
  Buf<double> zero;   // state placeholder of the mass mode
  Buf<int32_t> row_order, contrib, contrib_pos;
  Buf<int64_t> contrib_ptr;
  Buf<uint16_t> pos;
  Buf<double> geo_N, geo_dN, ref_tab, qwts, fn_c, elem_jac, elem_res;
  Buf<uint8_t> fn_op;
  Buf<int8_t> orient;
  std::vector<std::unique_ptr<SideDev>> sides;
};

const GenDeviceKernels* gen_find_device(const std::string& physics, int dim, int order, int nq, int nqs) {
  for (auto& k : device_table())
    if (physics == k.info.physics && dim == k.info.dim && order == k.info.order && nq == k.info.nq && (nqs == 0 || nqs == k.info.nqs)) return &k;
  return nullptr;
}

static GenEmulatorLookup g_emulator = nullptr;
void gen_set_emulator(GenEmulatorLookup fn) { g_emulator = fn; }
const GenHostKernels* gen_find_host(const std::string& physics, int dim, int order, int nq, int nqs) {
  return g_emulator ? static_cast<const GenHostKernels*>(g_emulator(physics.c_str(), dim, order, nq, nqs)) : nullptr;
}
std::string gen_supported_list() {
  std::string s;
  for (auto& k : device_table()) {
    if (!s.empty()) s += "; ";
    s += std::string(k.info.physics) + " dim " + std::to_string(k.info.dim) + " order " + std::to_string(k.info.order) + " nqp " + std::to_string(k.info.nq);
  }
  return s;
}

void gen_free(GeneralPlanDev* D) { delete D; }

GeneralPlanDev* gen_upload(const GeneralPlanHost& H, const MeshGraph& m, size_t* tot, std::string& err) {
  std::unique_ptr<GeneralPlanDev> D(new GeneralPlanDev());
  const size_t N = (size_t)H.info.N;
  // contributions by scratch slot and by position-table row (the host plan lists them by instance)
  // Tensor-core kernels index the scratch variable-major, (v, i) -> v * card + i (GenParams::var_major): kperm maps the
  // element-local dof to that index and the contribution list / position tables are permuted to match.
  std::vector<int> kperm(N);
  for (size_t i = 0; i < N; ++i) kperm[i] = (int)i;
  if (H.use_tensor)
    for (int v = 0; v < H.info.nvars; ++v)
      for (int i = 0; i < H.info.card[0]; ++i) kperm[(size_t)H.off[v][i]] = v * H.info.card[0] + i;
  std::vector<int32_t> cslot(H.contrib.size()), cpos(H.contrib.size());
  for (size_t p = 0; p < H.contrib.size(); ++p) {
    const int64_t inst = H.contrib[p] / (int64_t)N, i = kperm[(size_t)(H.contrib[p] % (int64_t)N)];
    cslot[p] = (int32_t)(gen_slot(H, inst) * (int64_t)N + i);
    cpos[p] = (int32_t)((int64_t)H.pos_id[(size_t)inst] * (int64_t)N + i);
  }
  std::vector<uint16_t> pos_dev;
  if (H.use_tensor) {
    pos_dev.resize(H.pos_tab.size());
    const size_t ntab = H.pos_tab.size() / (N * N);
    for (size_t t = 0; t < ntab; ++t)
      for (size_t r = 0; r < N; ++r)
        for (size_t c = 0; c < N; ++c) pos_dev[(t * N + (size_t)kperm[r]) * N + (size_t)kperm[c]] = H.pos_tab[(t * N + r) * N + c];
  }
  const std::vector<uint16_t>& pos_up = H.use_tensor ? pos_dev : H.pos_tab;
  const size_t n_scratch = (size_t)gen_scratch_instances(H);
  bool ok = D->row_order.upload(H.row_order, tot, err) && D->contrib.upload(cslot, tot, err) && D->contrib_pos.upload(cpos, tot, err) && D->contrib_ptr.upload(H.contrib_ptr, tot, err) &&
            D->pos.upload(pos_up, tot, err) && D->geo_N.upload(H.geo_N, tot, err) && D->geo_dN.upload(H.geo_dN, tot, err) &&
            D->ref_tab.upload(H.ref_tab, tot, err) && D->qwts.upload(H.qwts, tot, err) && D->fn_c.upload(H.fn_c, tot, err) &&
            D->fn_op.upload(H.fn_op, tot, err) && D->elem_jac.alloc(n_scratch * N * N, tot, err) && D->elem_res.alloc(n_scratch * N, tot, err);
  if (ok && !m.orient.empty()) ok = D->orient.upload(m.orient, tot, err);
  for (auto& s : H.sides) {
    if (!ok) break;
    std::unique_ptr<SideDev> sd(new SideDev());
    ok = sd->items.upload(s.items, tot, err) && sd->geo_N.upload(s.geo_N, tot, err) && sd->geo_dN.upload(s.geo_dN, tot, err) &&
         sd->ref_tab.upload(s.ref_tab, tot, err) && sd->qwts.upload(s.qwts, tot, err);
    D->sides.push_back(std::move(sd));
  }
  if (!ok) return nullptr;
  return D.release();
}

const char* gen_run(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                    const int32_t* conn, const int32_t* lids, const GraphDev& G, const OutDev& O, const double* sol, const TimeDev& td,
                    bool volume, bool boundary, void* stream, GenLaunchStats* stats, int pull_mass_mode, const double* mass_wts, bool adjoint = false);

static int pick_epb(const GenKernelInfo& I, bool tensor, bool side, int64_t n_items, int epb_override) {
  // as many elements per CTA as the launch bounds allow, while MINB CTAs still fit the SM's shared memory
  const int tpe = I.tpe;
  const int sd = tensor ? (side ? I.tc_smem_doubles_side : I.tc_smem_doubles_volume) : (side ? I.smem_doubles_side : I.smem_doubles_volume);
  const int max_threads = tensor ? I.tc_max_threads : I.max_threads;
  const size_t smem_cap = (size_t)(220 * 1024) / (size_t)std::max(1, tensor ? std::max(2, I.tc_min_blocks) : I.min_blocks);
  int epb = std::max(1, max_threads / tpe);
  if (tensor) epb = 16;   // tensor-core kernels: any thread count serves any element count (work items are strided over the CTA)
  if (epb_override > 0) epb = epb_override;
  while (epb > 1 && ((size_t)epb * sd * 8 > smem_cap || (!tensor && epb * tpe > max_threads))) --epb;
  if (n_items < epb) epb = (int)std::max<int64_t>(1, n_items);
  return epb;
}

const char* gen_assemble(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                         const int32_t* conn, const int32_t* lids, const GraphDev& G, const OutDev& O, const double* sol, const TimeDev& td,
                         bool volume, bool boundary, void* stream, GenLaunchStats* stats, bool adjoint) {
  return gen_run(D, H, kd, vx, vy, vz, conn, lids, G, O, sol, td, volume, boundary, stream, stats, 0, nullptr, adjoint);
}

const char* gen_assemble_mass(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                              const int32_t* conn, const int32_t* lids, const GraphDev& G, const double* mass_wts, bool lump, bool accumulate,
                              double* mass, double* diag, void* stream, GenLaunchStats* stats) {
  // the mass matrix does not depend on the state: any valid vector serves as `sol` (the element residual scratch is one)
  TimeDev td;
  std::memset(&td, 0, sizeof(td));
  td.alpha_u = 1.0; td.seed_u = 1.0; td.deltat = 1.0;
  OutDev O;
  O.jac = mass; O.res = diag; O.accumulate = accumulate ? 1 : 0; O.diag_one = 1;
  if (!D->zero.p) {
    std::string err;
    if (!D->zero.alloc((size_t)H.n_rows, nullptr, err)) return "general mass: cannot allocate the state placeholder";
    cudaMemset(D->zero.p, 0, (size_t)H.n_rows * sizeof(double));
  }
  return gen_run(D, H, kd, vx, vy, vz, conn, lids, G, O, D->zero.p, td, true, false, stream, stats, lump ? 2 : 1, mass_wts);
}

const char* gen_apply_mass(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                           const int32_t* conn, const int32_t* lids, const GraphDev& G, const double* mass_wts, bool accumulate, const double* x, double* y,
                           void* stream, GenLaunchStats* stats) {
  TimeDev td;
  std::memset(&td, 0, sizeof(td));
  td.alpha_u = 1.0; td.seed_u = 1.0; td.deltat = 1.0;
  OutDev O;
  O.jac = nullptr; O.res = y; O.accumulate = accumulate ? 1 : 0; O.diag_one = 1;
  return gen_run(D, H, kd, vx, vy, vz, conn, lids, G, O, x, td, true, false, stream, stats, 3, mass_wts);
}

const char* gen_project_initial(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                                const int32_t* conn, const int32_t* lids, const GraphDev& G, double time, bool accumulate, double* rhs, void* stream,
                                GenLaunchStats* stats) {
  TimeDev td;
  std::memset(&td, 0, sizeof(td));
  td.alpha_u = 1.0; td.seed_u = 1.0; td.deltat = 1.0; td.time = time;
  OutDev O;
  O.jac = nullptr; O.res = rhs; O.accumulate = accumulate ? 1 : 0; O.diag_one = 1;
  const double ones[GEN_MAXVARS] = {1.0, 1.0, 1.0, 1.0, 1.0};
  return gen_run(D, H, kd, vx, vy, vz, conn, lids, G, O, rhs /* state is not read in this mode: any valid vector */, td, true, false, stream, stats, 4, ones);
}

static const char* gen_run_impl_marker = nullptr;
const char* gen_run(GeneralPlanDev* D, const GeneralPlanHost& H, const GenDeviceKernels* kd, const double* vx, const double* vy, const double* vz,
                    const int32_t* conn, const int32_t* lids, const GraphDev& G, const OutDev& O, const double* sol, const TimeDev& td,
                    bool volume, bool boundary, void* stream, GenLaunchStats* stats, int pull_mass_mode, const double* mass_wts, bool adjoint) {
  (void)gen_run_impl_marker;
  const GenKernelInfo& I = H.info;
  GenParams P;
  std::memset(&P, 0, sizeof(P));
  P.vx = vx; P.vy = vy; P.vz = vz; P.conn = conn; P.lids = lids; P.orient = D->orient.n ? D->orient.p : nullptr;
  P.sol = sol; P.td = td;
  P.var_major = H.use_tensor ? 1 : 0;
  P.adjoint = adjoint ? 1 : 0;
  const bool initial = (pull_mass_mode == 4);   // projection of the initial conditions: the pull is the one of applyMassMatrixFree
  if (initial) pull_mass_mode = 3;
  if (pull_mass_mode) { P.mass_mode = initial ? 2 : 1; for (int v = 0; v < I.nvars; ++v) P.mass_wts[v] = mass_wts[v]; }
  std::memcpy(P.off, H.off, sizeof(P.off));
  std::memcpy(P.fn, initial ? H.init_fn : H.fn, sizeof(P.fn));
  auto mark_state = [&]() { P.fn_state = 0; if (!pull_mass_mode) for (int f = 0; f < GEN_MAXFN; ++f) if ((P.fn[f].pad & 1) && !P.fn[f].is_const) P.fn_state = 1; };
  mark_state();
  P.fn_op = D->fn_op.p; P.fn_c = D->fn_c.p; P.opt = H.opt;
  if (adjoint && std::string(I.physics).find('+') != std::string::npos) return "adjoint assembly of a two-module block is not built (thermal's sf = 1 and the other module's form_param share one option slot)";
  if (adjoint && std::string(I.physics) == "thermal") P.opt.form_param = 1.0;   // thermal.cpp:197-201, 292-296: sf = 1 when wkset->isAdjoint
  for (int v = 0; v < GEN_MAXVARS; ++v) { P.bc_type[v] = 0; P.bc_fn[v] = -1; }
  P.elem_jac = (pull_mass_mode == 3) ? nullptr : ((O.jac || pull_mass_mode) ? D->elem_jac.p : nullptr);
  P.elem_res = (pull_mass_mode == 3 || (O.res && !pull_mass_mode)) ? D->elem_res.p : nullptr;
  int launches = 0;
  auto run_elements = [&](bool side, int64_t n_items) -> const char* {
    if (n_items <= 0) return nullptr;
    const bool tensor = H.use_tensor && P.elem_jac != nullptr;   // residual-only calls have no use for the wider CTAs
    const int epb = pick_epb(I, tensor, side, n_items, H.epb_override);
    P.epb = epb;
    const int tpe = I.tpe;
    int threads = ((epb * tpe + 31) / 32) * 32;
    threads = std::max(32, std::min(I.max_threads, threads));
    if (tensor) threads = std::max(64, std::min(I.tc_max_threads, 32 * epb * I.nvars * I.nvars * (I.card[0] > 24 ? 2 : 1)));   // one warp per block of the contraction (S4m), two warps at least
    const size_t smem = (size_t)epb * (tensor ? (side ? I.tc_smem_doubles_side : I.tc_smem_doubles_volume) : (side ? I.smem_doubles_side : I.smem_doubles_volume)) * sizeof(double);
    const int64_t nblocks = (n_items + epb - 1) / epb;
    ++launches;
    const int st = P.fn_state ? 1 : 0;
    return tensor ? kd->launch_tc[st](side, P, (int)nblocks, threads, smem, stream) : kd->launch[st](side, P, (int)nblocks, threads, smem, stream);
  };
  auto run_pull = [&](int64_t row_begin, int64_t row_end) -> const char* {
    if (row_end <= row_begin) return nullptr;
    PullParams Q;
    Q.row_order = D->row_order.p; Q.contrib_ptr = D->contrib_ptr.p; Q.contrib = D->contrib.p; Q.contrib_pos = D->contrib_pos.p; Q.pos = D->pos.p;
    Q.elem_jac = D->elem_jac.p; Q.elem_res = D->elem_res.p; Q.G = G; Q.O = O;
    Q.row_begin = row_begin; Q.row_end = row_end; Q.n_owned = H.n_owned; Q.N = I.N; Q.max_row_len = std::max(1, H.max_row_len);
    Q.mass_mode = pull_mass_mode; Q.mass_inst_end = H.scratch_cap;
    Q.lump = (H.lump_mass && !pull_mass_mode) ? 1 : 0;
    const int Gs = I.N <= 8 ? 8 : (I.N <= 16 ? 16 : 32);
    const int npl = (I.N + Gs - 1) / Gs;
    if (npl > 3) return "general pull: more than 96 dofs per element";
    const int threads = 256, rows_per_block = threads / Gs;
    const size_t smem = (size_t)rows_per_block * Q.max_row_len * sizeof(double);
    const int64_t nblocks = (row_end - row_begin + rows_per_block - 1) / rows_per_block;
    ++launches;
    typedef void (*PullFn)(PullParams);
    const PullFn fn = Gs == 8 ? (PullFn)gen_pull_kernel<8, 1> : Gs == 16 ? (PullFn)gen_pull_kernel<16, 1>
                    : npl == 1 ? (PullFn)gen_pull_kernel<32, 1> : npl == 2 ? (PullFn)gen_pull_kernel<32, 2> : (PullFn)gen_pull_kernel<32, 3>;
    if (smem > 48 * 1024) {
      const cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cudaGetErrorString(e);
    }
    fn<<<(int)nblocks, threads, smem, (cudaStream_t)stream>>>(Q);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
  };
  const bool any_side = boundary && std::any_of(H.sides.begin(), H.sides.end(), [](const GenSideFamily& s) { return s.active && !s.items.empty(); });
  // instances that are not computed in this call must read as zero: clear the scratch ranges that are skipped
  if (!volume) {   // the whole ring reads as zero
    if (P.elem_jac) cudaMemsetAsync(D->elem_jac.p, 0, (size_t)H.scratch_cap * I.N * I.N * sizeof(double), (cudaStream_t)stream);
    if (P.elem_res) cudaMemsetAsync(D->elem_res.p, 0, (size_t)H.scratch_cap * I.N * sizeof(double), (cudaStream_t)stream);
  }
  if (!any_side && H.n_inst > H.n_elem) {
    if (P.elem_jac) cudaMemsetAsync(D->elem_jac.p + (size_t)H.scratch_cap * I.N * I.N, 0, (size_t)(H.n_inst - H.n_elem) * I.N * I.N * sizeof(double), (cudaStream_t)stream);
    if (P.elem_res) cudaMemsetAsync(D->elem_res.p + (size_t)H.scratch_cap * I.N, 0, (size_t)(H.n_inst - H.n_elem) * I.N * sizeof(double), (cudaStream_t)stream);
  }
  const int nb = (int)H.batches.size();
  for (int b = 0; b < nb; ++b) {
    const GenBatch& B = H.batches[(size_t)b];
    if (volume) {
      P.items = nullptr; P.item_begin = B.elem_begin; P.item_end = B.elem_end; P.inst_base = B.elem_begin % H.scratch_cap;   // batches do not wrap: the ring is a multiple of the batch size
      std::memcpy(P.fn, initial ? H.init_fn : H.fn, sizeof(P.fn));
      mark_state();
      for (int v = 0; v < GEN_MAXVARS; ++v) { P.bc_type[v] = 0; P.bc_fn[v] = -1; }
      P.geo_N = D->geo_N.p; P.geo_dN = D->geo_dN.p; P.ref_tab = D->ref_tab.p; P.qwts = D->qwts.p;
      if (const char* e = run_elements(false, B.elem_end - B.elem_begin)) return e;
    }
    if (b == nb - 1 && any_side) {
      for (size_t s = 0; s < H.sides.size(); ++s) {
        const GenSideFamily& S = H.sides[s];
        if (!S.active || S.items.empty()) continue;
        const SideDev& sd = *D->sides[s];
        P.items = sd.items.p; P.item_begin = 0; P.item_end = (int64_t)S.items.size(); P.inst_base = H.scratch_cap + (S.inst_base - H.n_elem);
        P.geo_N = sd.geo_N.p; P.geo_dN = sd.geo_dN.p; P.ref_tab = sd.ref_tab.p; P.qwts = sd.qwts.p;
        for (int d = 0; d < 3; ++d) { P.tan_u[d] = S.tan_u[d]; P.tan_v[d] = S.tan_v[d]; }
        for (int v = 0; v < GEN_MAXVARS; ++v) { P.bc_type[v] = S.bc_type[v]; P.bc_fn[v] = S.bc_fn[v]; }
        std::memcpy(P.fn, S.fn, sizeof(P.fn));
        mark_state();
        if (const char* e = run_elements(true, (int64_t)S.items.size())) return e;
      }
    }
    if (const char* e = run_pull(B.row_begin, B.row_end)) return e;
  }
  if (stats) stats->launches = launches;
  return nullptr;
}

}  // namespace mrhyde_b200
