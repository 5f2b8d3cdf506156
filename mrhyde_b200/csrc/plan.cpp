// Host-side plan construction (see plan.hpp).
#include "plan.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <unordered_map>

namespace mrhyde_b200 {

namespace {

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
inline uint64_t bits_of(double d) { if (d == 0.0) d = 0.0; uint64_t u; std::memcpy(&u, &d, 8); return u; }

int host_threads() {
  int nt = (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 16) nt = 16;
  return nt;
}

template <class F>
void parallel_for(int64_t n, F&& f) {
  int nt = host_threads();
  if (n < 4 * nt) nt = 1;
  std::atomic<int64_t> next(0);
  std::atomic<bool> failed(false);
  std::string what;
  std::mutex mu;
  auto worker = [&](int tid) {
    try {
      for (;;) {
        const int64_t i = next.fetch_add(1);
        if (i >= n || failed.load()) break;
        f(i, tid);
      }
    } catch (const std::exception& e) {
      std::lock_guard<std::mutex> lock(mu);
      if (!failed.exchange(true)) what = e.what();
    }
  };
  if (nt == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  if (failed.load()) throw std::runtime_error(what);
}

}  // namespace

void MeshGraph::set_elem_nodes(int64_t n_elem, const double* elem_nodes) {
  nelem = n_elem;
  const int64_t n = n_elem * nverts;
  uint64_t cap = 16;
  while (cap < (uint64_t)n * 2 + 16) cap <<= 1;
  std::vector<int32_t> table(cap, -1);
  for (int d = 0; d < 3; ++d) vcoord[d].clear();
  conn.assign((size_t)n, 0);
  for (int64_t k = 0; k < n; ++k) {
    double c[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) c[d] = elem_nodes[k * dim + d];
    uint64_t h = mix64(bits_of(c[0]) ^ mix64(bits_of(c[1]) + 0x9e3779b97f4a7c15ULL) ^ mix64(bits_of(c[2]) + 0x7f4a7c159e3779b9ULL));
    uint64_t slot = h & (cap - 1);
    for (;;) {
      const int32_t v = table[slot];
      if (v < 0) {
        const int32_t id = (int32_t)vcoord[0].size();
        for (int d = 0; d < 3; ++d) vcoord[d].push_back(c[d]);
        table[slot] = id;
        conn[(size_t)k] = id;
        break;
      }
      if (vcoord[0][v] == c[0] && vcoord[1][v] == c[1] && vcoord[2][v] == c[2]) { conn[(size_t)k] = v; break; }
      slot = (slot + 1) & (cap - 1);
    }
  }
  nvert = (int64_t)vcoord[0].size();
}

void MeshGraph::classify_cells() {
  eclass.assign((size_t)nelem, 0);
  const double tol = 2e-14;
  for (int64_t e = 0; e < nelem; ++e) {
    const int32_t* c = &conn[(size_t)e * nverts];
    double X[8][3];
    for (int n = 0; n < nverts; ++n) for (int d = 0; d < 3; ++d) X[n][d] = vcoord[d][c[n]];
    double worst = 0.0, emin = 1e300;
    auto edge_len = [&](int a, int b) {
      double s = 0; for (int d = 0; d < dim; ++d) s += (X[a][d] - X[b][d]) * (X[a][d] - X[b][d]);
      return std::sqrt(s);
    };
    if (dim == 2) {
      emin = std::min(edge_len(1, 0), edge_len(3, 0));
      for (int d = 0; d < 2; ++d) worst = std::max(worst, std::fabs(X[2][d] - (X[1][d] + X[3][d] - X[0][d])));
    } else {
      emin = std::min(edge_len(1, 0), std::min(edge_len(3, 0), edge_len(4, 0)));
      for (int d = 0; d < 3; ++d) {
        worst = std::max(worst, std::fabs(X[2][d] - (X[1][d] + X[3][d] - X[0][d])));
        worst = std::max(worst, std::fabs(X[5][d] - (X[1][d] + X[4][d] - X[0][d])));
        worst = std::max(worst, std::fabs(X[7][d] - (X[3][d] + X[4][d] - X[0][d])));
        worst = std::max(worst, std::fabs(X[6][d] - (X[1][d] + X[3][d] + X[4][d] - 2.0 * X[0][d])));
      }
    }
    if (!(worst <= tol * emin)) continue;
    // edge vectors from vertex 0 along xi, eta, zeta (Shards order: 1, 3, 4); "diagonal" must be exact so that the
    // box path drops only terms that are exactly zero
    static const int nb[3] = {1, 3, 4};
    bool diag = true;
    for (int a = 0; a < dim; ++a)
      for (int d = 0; d < dim; ++d)
        if (d != a && X[nb[a]][d] - X[0][d] != 0.0) diag = false;
    eclass[(size_t)e] = diag ? 2 : 1;
  }
}

namespace {

struct ChainWork {  // per-chain scratch produced in parallel, concatenated afterwards
  std::vector<StepRec> steps;       // elem_begin / batch_begin local to this chain
  std::vector<int32_t> elems;
  std::vector<BatchRec> batches;    // row_begin local to this chain
  std::vector<RowRec> rows;
};

}  // namespace

// MRHYDE_B200_PLAN_TIMING=1: wall time of the plan builder's phases on stderr
struct PhaseTimer {
  bool on = std::getenv("MRHYDE_B200_PLAN_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[mrhyde_b200 plan] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

void build_chain_plan(const MeshGraph& m, const std::vector<uint16_t>& kmap, const std::vector<uint16_t>& rmap,
                      int stage_len, const ChainOptions& opt, ChainPlan& out) {
  const int64_t ne = m.nelem, nr = m.nrows;
  const int nd = m.ndof, nv = m.nverts, dim = m.dim;
  if (ne <= 0 || nr <= 0) throw std::runtime_error("plan: empty mesh or graph");
  const int axis = (opt.sweep_axis >= 0 && opt.sweep_axis < dim) ? opt.sweep_axis : dim - 1;

  PhaseTimer phase;
  // ---- row -> (element, local dof) adjacency, elements ascending
  std::vector<int64_t> r2e_ptr((size_t)nr + 1, 0);
  for (int64_t k = 0; k < ne * nd; ++k) {
    const int32_t r = m.lids[(size_t)k];
    if (r < 0 || r >= nr) throw std::runtime_error("plan: LID out of range of the graph");
    ++r2e_ptr[(size_t)r + 1];
  }
  for (int64_t r = 0; r < nr; ++r) r2e_ptr[(size_t)r + 1] += r2e_ptr[(size_t)r];
  std::vector<int32_t> r2e_elem((size_t)(ne * nd));
  std::vector<uint8_t> r2e_dof((size_t)(ne * nd));
  {
    std::vector<int64_t> fill(r2e_ptr.begin(), r2e_ptr.end() - 1);
    for (int64_t e = 0; e < ne; ++e)
      for (int i = 0; i < nd; ++i) {
        const int32_t r = m.lids[(size_t)e * nd + i];
        const int64_t p = fill[(size_t)r]++;
        r2e_elem[(size_t)p] = (int32_t)e;
        r2e_dof[(size_t)p] = (uint8_t)i;
      }
  }

  phase.lap("row adjacency");
  // ---- element centroids and lowest coordinate along the sweep axis
  std::vector<double> cen[3], elo((size_t)ne);
  for (int d = 0; d < 3; ++d) cen[d].assign((size_t)ne, 0.0);
  double amin = 1e300, amax = -1e300;
  for (int64_t e = 0; e < ne; ++e) {
    double lo = 1e300;
    for (int n = 0; n < nv; ++n) {
      const int32_t v = m.conn[(size_t)e * nv + n];
      for (int d = 0; d < dim; ++d) cen[d][(size_t)e] += m.vcoord[d][(size_t)v];
      lo = std::min(lo, m.vcoord[axis][(size_t)v]);
      amax = std::max(amax, m.vcoord[axis][(size_t)v]);
    }
    for (int d = 0; d < dim; ++d) cen[d][(size_t)e] /= nv;
    elo[(size_t)e] = lo;
    amin = std::min(amin, lo);
  }
  const double atol = 1e-9 * std::max(amax - amin, 1e-300);

  phase.lap("centroids");
  // ---- levels: breadth-first sweep over "shares a dof" adjacency, seeded at the low face of the sweep axis
  std::vector<int32_t> level((size_t)ne, -1);
  {
    std::vector<uint8_t> row_done((size_t)nr, 0);
    std::vector<int32_t> frontier, next;
    int64_t visited = 0;
    double seed_lo = amin;
    while (visited < ne) {
      frontier.clear();
      for (int64_t e = 0; e < ne; ++e)
        if (level[(size_t)e] < 0 && elo[(size_t)e] <= seed_lo + atol) { level[(size_t)e] = 0; frontier.push_back((int32_t)e); }
      visited += (int64_t)frontier.size();
      int32_t lv = 0;
      while (!frontier.empty()) {
        next.clear();
        for (int32_t e : frontier)
          for (int i = 0; i < nd; ++i) {
            const int32_t r = m.lids[(size_t)e * nd + i];
            if (row_done[(size_t)r]) continue;
            row_done[(size_t)r] = 1;
            for (int64_t p = r2e_ptr[(size_t)r]; p < r2e_ptr[(size_t)r + 1]; ++p) {
              const int32_t f = r2e_elem[(size_t)p];
              if (level[(size_t)f] < 0) { level[(size_t)f] = lv + 1; next.push_back(f); }
            }
          }
        visited += (int64_t)next.size();
        frontier.swap(next);
        ++lv;
      }
      if (visited < ne) {  // disconnected remainder: restart from its own low face
        seed_lo = 1e300;
        for (int64_t e = 0; e < ne; ++e) if (level[(size_t)e] < 0) seed_lo = std::min(seed_lo, elo[(size_t)e]);
      }
    }
  }
  int32_t nlevels = 0;
  for (int64_t e = 0; e < ne; ++e) nlevels = std::max(nlevels, level[(size_t)e] + 1);

  phase.lap("levels");
  // axis along which consecutive element ids are displaced most often (x on the reference's inline meshes)
  int fast_axis = 0;
  {
    int64_t votes[3] = {0, 0, 0};
    for (int64_t e = 0; e + 1 < ne; ++e) {
      int best = 0;
      double bd = -1.0;
      for (int d = 0; d < dim; ++d) { const double v = std::fabs(cen[d][(size_t)e + 1] - cen[d][(size_t)e]); if (v > bd) { bd = v; best = d; } }
      ++votes[best];
    }
    for (int d = 1; d < dim; ++d) if (votes[d] > votes[fast_axis]) fast_axis = d;
  }
  const int cap_limit = (int)std::min<size_t>(opt.smem_budget / 2 / ((size_t)stage_len * 8), 256);   // one thread per element, <= 256 threads
  if (cap_limit < 1) throw std::runtime_error("plan: a single element does not fit the shared-memory ring");
  int column_elems = std::max(1, std::min(opt.column_elems, cap_limit));

  for (int attempt = 0;; ++attempt) {
    out = ChainPlan();
    out.stage_len = stage_len;
    out.n_levels = nlevels;
    // ---- columns: recursive coordinate bisection of the projected centroids, cuts only between distinct coordinates
    std::vector<int32_t> col((size_t)ne, 0);
    int32_t ncol = 0;
    {
      std::vector<int32_t> idx((size_t)ne);
      for (int64_t e = 0; e < ne; ++e) idx[(size_t)e] = (int32_t)e;
      struct Range { int64_t b, e; };
      std::vector<Range> stack{{0, ne}};
      std::vector<double> keys;
      const int64_t target = (int64_t)column_elems * nlevels;
      while (!stack.empty()) {
        const Range rg = stack.back();
        stack.pop_back();
        const int64_t n = rg.e - rg.b;
        bool split = false;
        if (n > target && dim > 1) {
          // candidate axes ordered by centroid spread
          int axes[2], na = 0;
          double spread[2];
          for (int d = 0; d < dim; ++d) {
            if (d == axis) continue;
            double lo = 1e300, hi = -1e300;
            for (int64_t k = rg.b; k < rg.e; ++k) { const double c = cen[d][(size_t)idx[(size_t)k]]; lo = std::min(lo, c); hi = std::max(hi, c); }
            axes[na] = d; spread[na] = hi - lo; ++na;
          }
          if (na == 2 && spread[1] > spread[0] * (1.0 + 1e-9)) { std::swap(axes[0], axes[1]); std::swap(spread[0], spread[1]); }
          // equal spreads: cut across the slower axis so that columns stay long along the axis element ids run fastest in --
          // rows of a batch then sit next to each other in the ring (conflict-free shared-memory reads) and in the CSR array
          else if (na == 2 && !(spread[0] > spread[1] * (1.0 + 1e-9)) && axes[0] == fast_axis) { std::swap(axes[0], axes[1]); std::swap(spread[0], spread[1]); }
          for (int a = 0; a < na && !split; ++a) {
            const int d = axes[a];
            if (!(spread[a] > 0.0)) continue;
            // the cut sits between two distinct coordinate values, so the two halves are {c < v} and {c >= v} for the value v right of
            // the cut: sorting the coordinate VALUES finds v, a stable partition of the ids applies it (the ids of a range stay
            // ascending; an indirect sort of the ids gave the same halves at several times the cost)
            keys.resize((size_t)n);
            for (int64_t k = 0; k < n; ++k) keys[(size_t)k] = cen[d][(size_t)idx[(size_t)(rg.b + k)]];
            std::sort(keys.begin(), keys.end());
            const double tol = 1e-9 * spread[a];
            const int64_t mid = n / 2;
            int64_t cut = -1;
            for (int64_t off = 0; off < n; ++off) {  // nearest position to the median where the coordinate changes
              const int64_t c1 = mid + off, c2 = mid - off;
              if (c1 > 0 && c1 < n && keys[(size_t)c1] - keys[(size_t)c1 - 1] > tol) { cut = c1; break; }
              if (c2 > 0 && c2 < n && keys[(size_t)c2] - keys[(size_t)c2 - 1] > tol) { cut = c2; break; }
            }
            if (cut > 0 && cut < n) {
              const double v = keys[(size_t)cut];
              std::stable_partition(idx.begin() + rg.b, idx.begin() + rg.e, [&](int32_t x) { return cen[d][(size_t)x] < v; });
              stack.push_back({rg.b + cut, rg.e}); stack.push_back({rg.b, rg.b + cut}); split = true;
            }
          }
        }
        if (!split) {
          for (int64_t k = rg.b; k < rg.e; ++k) col[(size_t)idx[(size_t)k]] = ncol;
          ++ncol;
        }
      }
    }
    out.n_columns = ncol;
    // segments along the sweep: enough chains to fill the device, preferring counts that make whole waves of
    // resident CTAs; every extra segment costs one recomputed level per column
    int32_t nseg = 1;
    int cta_slots = opt.cta_slots;
    if (cta_slots <= 0) {
      // resident CTAs per SM from the shared memory of one CTA: the ring holds one step of a column plus its halo ring, estimated
      // here from the column's cross-section (the exact capacity is only known after the chains are laid out)
      const double side = std::sqrt((double)column_elems);
      const double cap_est = dim == 3 ? (double)column_elems + 2.0 * side + 1.0 : (dim == 2 ? (double)column_elems + 1.0 : 1.0);
      const int threads_est = std::max(128, std::min(256, (((int)cap_est + 31) / 32) * 32));
      const size_t stage = (size_t)(opt.ring_stage_len > 0 ? opt.ring_stage_len : stage_len);
      const size_t smem_est = (size_t)(2.0 * cap_est) * stage * 8 + (size_t)(threads_est / 32) * opt.warp_buffer_bytes + 1024;
      const int blocks = std::max(1, std::min(std::min(opt.max_blocks_per_sm, 2048 / threads_est), (int)((228 * 1024) / smem_est)));
      cta_slots = std::max(1, opt.n_sm) * blocks;
    }
    {
      const int32_t smax = std::max(1, nlevels / std::max(1, opt.min_segment_levels));
      const int32_t smin = std::max(1, std::min(smax, (int32_t)((opt.min_chains + ncol - 1) / ncol)));
      double best = -1.0;
      for (int32_t sg = smin; sg <= std::min(smax, 4 * smin + 4); ++sg) {
        const double waves = (double)ncol * sg / std::max(1, cta_slots);
        const double fill = waves / std::ceil(waves);
        const double work = (double)nlevels / (double)(nlevels + sg - 1);
        const double score = fill * work;
        if (score > best + 1e-12) { best = score; nseg = sg; }
      }
    }
    out.n_segments = nseg;
    auto seg_of = [&](int32_t lv) { return (int32_t)(((int64_t)lv * nseg) / nlevels); };
    const int32_t nchains = ncol * nseg;
    out.n_chains = nchains;

    phase.lap("columns + segments");
    // ---- row owner: the chain of the adjacent element with the largest (level, column); row completes at that level
    std::vector<int32_t> row_chain((size_t)nr, -1), row_level((size_t)nr, -1);
    std::vector<int32_t> chain_row_ptr((size_t)nchains + 1, 0);
    for (int64_t r = 0; r < nr; ++r) {
      int32_t bl = -1, bc = -1;
      for (int64_t p = r2e_ptr[(size_t)r]; p < r2e_ptr[(size_t)r + 1]; ++p) {
        const int32_t e = r2e_elem[(size_t)p];
        const int32_t l = level[(size_t)e], c = col[(size_t)e];
        if (l > bl || (l == bl && c > bc)) { bl = l; bc = c; }
      }
      if (bl < 0) { out.orphan_rows.push_back((int32_t)r); continue; }
      row_chain[(size_t)r] = bc * nseg + seg_of(bl);
      row_level[(size_t)r] = bl;
      ++chain_row_ptr[(size_t)row_chain[(size_t)r] + 1];
    }
    // multi-rank plans: chains that complete ghost rows (rows another rank owns, local id >= nowned) get the lowest chain ids,
    // so that they can be launched first and their rows sent to the owner while the remaining chains are assembled
    out.n_early_chains = 0;
    if (m.nowned > 0 && m.nowned < nr) {
      std::vector<uint8_t> early((size_t)nchains, 0);
      for (int64_t r = m.nowned; r < nr; ++r) if (row_chain[(size_t)r] >= 0) early[(size_t)row_chain[(size_t)r]] = 1;
      std::vector<int32_t> newid((size_t)nchains, 0);
      int32_t k = 0;
      for (int32_t c = 0; c < nchains; ++c) if (early[(size_t)c]) newid[(size_t)c] = k++;
      out.n_early_chains = k;
      for (int32_t c = 0; c < nchains; ++c) if (!early[(size_t)c]) newid[(size_t)c] = k++;
      std::fill(chain_row_ptr.begin(), chain_row_ptr.end(), 0);
      for (int64_t r = 0; r < nr; ++r)
        if (row_chain[(size_t)r] >= 0) { row_chain[(size_t)r] = newid[(size_t)row_chain[(size_t)r]]; ++chain_row_ptr[(size_t)row_chain[(size_t)r] + 1]; }
    }
    for (int32_t c = 0; c < nchains; ++c) chain_row_ptr[(size_t)c + 1] += chain_row_ptr[(size_t)c];
    std::vector<int32_t> chain_rows((size_t)chain_row_ptr[(size_t)nchains]);
    {
      std::vector<int32_t> fill(chain_row_ptr.begin(), chain_row_ptr.end() - 1);
      for (int64_t r = 0; r < nr; ++r) if (row_chain[(size_t)r] >= 0) chain_rows[(size_t)fill[(size_t)row_chain[(size_t)r]]++] = (int32_t)r;
    }

    phase.lap("row owners");
    // ---- pass 1: per-chain element lists by level (own column + halo ring), ring capacity
    const int nthreads = host_threads();
    std::vector<ChainWork> work((size_t)nchains);
    std::vector<std::vector<int32_t>> stamp((size_t)nthreads);
    std::atomic<int32_t> max_step_elems(0);
    parallel_for(nchains, [&](int64_t c, int tid) {
      auto& st = stamp[(size_t)tid];
      if (st.empty()) st.assign((size_t)ne, -1);
      ChainWork& W = work[(size_t)c];
      std::vector<int32_t> need;
      for (int32_t k = chain_row_ptr[(size_t)c]; k < chain_row_ptr[(size_t)c + 1]; ++k) {
        const int32_t r = chain_rows[(size_t)k];
        for (int64_t q = r2e_ptr[(size_t)r]; q < r2e_ptr[(size_t)r + 1]; ++q) {
          const int32_t e = r2e_elem[(size_t)q];
          if (st[(size_t)e] != (int32_t)c) { st[(size_t)e] = (int32_t)c; need.push_back(e); }
        }
      }
      std::sort(need.begin(), need.end(), [&](int32_t a, int32_t b) {
        return level[(size_t)a] < level[(size_t)b] || (level[(size_t)a] == level[(size_t)b] && a < b);
      });
      W.elems = need;
      size_t k = 0;
      while (k < need.size()) {
        size_t k2 = k;
        while (k2 < need.size() && level[(size_t)need[k2]] == level[(size_t)need[k]]) ++k2;
        StepRec S;
        S.elem_begin = (int32_t)k; S.n_elem = (int32_t)(k2 - k); S.batch_begin = 0; S.n_batches = 0;
        W.steps.push_back(S);
        int32_t cur = max_step_elems.load();
        while (S.n_elem > cur && !max_step_elems.compare_exchange_weak(cur, S.n_elem)) {}
        k = k2;
      }
    });
    if (max_step_elems.load() > cap_limit) {
      if (column_elems == 1 || attempt > 40) throw std::runtime_error("plan: cannot fit one sweep step into the shared-memory ring");
      column_elems = std::max(1, (column_elems * 3) / 4);
      continue;
    }
    out.cap = std::max(1, max_step_elems.load());
    const uint32_t cap = (uint32_t)out.cap;
    const uint32_t slot_bytes = (uint32_t)out.slot_bytes();
    const bool metric_ok = nd <= 8 && cap <= 256;   // field widths of the metric source word

    phase.lap("pass 1 (element lists)");
    // ---- pass 2: rows of every step and their gather patterns
    std::unordered_map<std::string, int32_t> pattern_ids;
    std::mutex mu;
    struct Scratch {
      std::vector<int32_t> pos_cur, pos_prev;           // global element -> index in the step's list (stamped by tag)
      std::vector<int32_t> tag_cur, tag_prev;
      std::unordered_map<std::string, int32_t> cache;   // thread-local view of pattern_ids
      std::vector<std::vector<uint32_t>> slot_src;      // per CSR slot (+ residual): packed (slot_rel, le, t)
      std::vector<std::vector<uint8_t>> slot_ij;        // same order: local row i << 3 | local column j (metric words)
      std::string key;
    };
    std::vector<Scratch> scratch((size_t)nthreads);
    std::atomic<int32_t> max_rows_step(0), max_batches_step(0);
    int64_t tagbase = 0;
    std::vector<int64_t> chain_tag((size_t)nchains);
    for (int32_t c = 0; c < nchains; ++c) { chain_tag[(size_t)c] = tagbase; tagbase += (int64_t)work[(size_t)c].steps.size() + 2; }
    if (tagbase > 0x7fffffff) throw std::runtime_error("plan: too many sweep steps");
    struct StepRow { int32_t pid; RowRec rec; };
    parallel_for(nchains, [&](int64_t c, int tid) {
      Scratch& S = scratch[(size_t)tid];
      if (S.pos_cur.empty()) { S.pos_cur.assign((size_t)ne, 0); S.pos_prev.assign((size_t)ne, 0); S.tag_cur.assign((size_t)ne, -1); S.tag_prev.assign((size_t)ne, -1); }
      ChainWork& W = work[(size_t)c];
      // rows of this chain bucketed by completion level
      std::vector<int32_t> rows_sorted(chain_rows.begin() + chain_row_ptr[(size_t)c], chain_rows.begin() + chain_row_ptr[(size_t)c + 1]);
      std::sort(rows_sorted.begin(), rows_sorted.end(), [&](int32_t a, int32_t b) {
        return row_level[(size_t)a] < row_level[(size_t)b] || (row_level[(size_t)a] == row_level[(size_t)b] && a < b);
      });
      std::vector<StepRow> step_rows;
      size_t rk = 0;
      for (size_t s = 0; s < W.steps.size(); ++s) {
        StepRec& ST = W.steps[s];
        const int32_t lv = level[(size_t)W.elems[(size_t)ST.elem_begin]];
        const int32_t tag = (int32_t)(chain_tag[(size_t)c] + (int64_t)s);
        // previous step's map becomes pos_prev (swap roles), current step is stamped fresh
        S.pos_cur.swap(S.pos_prev); S.tag_cur.swap(S.tag_prev);
        for (int32_t l = 0; l < ST.n_elem; ++l) { const int32_t e = W.elems[(size_t)(ST.elem_begin + l)]; S.pos_cur[(size_t)e] = l; S.tag_cur[(size_t)e] = tag; }
        step_rows.clear();
        if (rk < rows_sorted.size() && row_level[(size_t)rows_sorted[rk]] < lv) throw std::runtime_error("plan: a row precedes its sweep step (internal error)");
        for (; rk < rows_sorted.size() && row_level[(size_t)rows_sorted[rk]] == lv; ++rk) {
          const int32_t r = rows_sorted[rk];
          const int64_t rs = m.rowptr[(size_t)r], re = m.rowptr[(size_t)r + 1];
          const int32_t len = (int32_t)(re - rs);
          if (len > 0xFFFE) throw std::runtime_error("plan: CSR row longer than 65534 entries");
          const int32_t* cols = &m.colind[(size_t)rs];
          StepRow SR;
          SR.rec.row = r; SR.rec.anchor = 0; SR.rec.aux = 0xFFFF; SR.rec.base = 0;
          if (m.fixed[(size_t)r]) {  // strong-Dirichlet row: nothing is gathered (scatter.hpp:208, 253)
            const int32_t* dg = std::lower_bound(cols, cols + len, r);
            SR.rec.aux = (dg != cols + len && *dg == r) ? (uint16_t)(dg - cols) : (uint16_t)0xFFFF;
            // a ghost copy of a fixed row contributes nothing: only the owner writes J(d,d) = 1, so the halo-summed
            // matrix equals the serial reference's (the MPI reference would add one per sharing rank)
            if (m.nowned > 0 && r >= m.nowned) SR.rec.aux = 0xFFFF;
            SR.pid = -1;
            step_rows.push_back(SR);
            continue;
          }
          if (S.slot_src.size() < (size_t)len + 1) { S.slot_src.resize((size_t)len + 1); S.slot_ij.resize((size_t)len + 1); }
          for (int32_t k = 0; k <= len; ++k) { S.slot_src[(size_t)k].clear(); S.slot_ij[(size_t)k].clear(); }
          uint32_t anchor = 0xFFFFFFFFu;
          for (int64_t q = r2e_ptr[(size_t)r]; q < r2e_ptr[(size_t)r + 1]; ++q) {
            const int32_t e = r2e_elem[(size_t)q];
            const int i = r2e_dof[(size_t)q];
            uint32_t rel, le;
            if (S.tag_cur[(size_t)e] == tag) { rel = 0; le = (uint32_t)S.pos_cur[(size_t)e]; }
            else if (s > 0 && S.tag_prev[(size_t)e] == tag - 1) { rel = 1; le = (uint32_t)S.pos_prev[(size_t)e]; }
            else throw std::runtime_error("plan: an element of row " + std::to_string(r) + " is outside the two-step window (internal error)");
            anchor = std::min(anchor, le);
            for (int j = 0; j < nd; ++j) {
              const int32_t cj = m.lids[(size_t)e * nd + j];
              const int32_t* it = std::lower_bound(cols, cols + len, cj);
              if (it == cols + len || *it != cj) throw std::runtime_error("plan: graph is missing an element coupling (row " + std::to_string(r) + ", col " + std::to_string(cj) + ")");
              S.slot_src[(size_t)(it - cols)].push_back((rel << 31) | (le << 12) | (uint32_t)kmap[(size_t)i * nd + j]);
              S.slot_ij[(size_t)(it - cols)].push_back((uint8_t)((i << 3) | j));
            }
            S.slot_src[(size_t)len].push_back((rel << 31) | (le << 12) | (uint32_t)rmap[(size_t)i]);
            S.slot_ij[(size_t)len].push_back((uint8_t)(i << 3));
          }
          // slot descriptors (parity 0), keyed for de-duplication
          std::string& key = S.key;
          key.clear();
          auto put = [&](uint32_t v) { key.append((const char*)&v, 4); };
          for (int32_t k = 0; k <= len; ++k) {
            const auto& src = S.slot_src[(size_t)k];
            if (src.size() > (size_t)SLOT_SRCS) throw std::runtime_error("plan: more than 8 elements contribute to one matrix entry");
            for (int z = 0; z < SLOT_SRCS; ++z) {
              if ((size_t)z < src.size()) {
                const uint32_t rel = src[(size_t)z] >> 31, le = (src[(size_t)z] >> 12) & 0x7FFFFu, t = src[(size_t)z] & 0xFFFu;
                put(rel * slot_bytes + (t * cap + (le - anchor)) * 8u);   // parity-0 offset; parity 1 flips the slot
              } else put(SRC_NONE);
            }
          }
          const size_t legacy_words = key.size() / 4;
          if (metric_ok)   // metric-ring words (kernel_abi.h), appended to the key: a pattern is one (desc, mdesc) pair
            for (int32_t k = 0; k <= len; ++k) {
              const auto& src = S.slot_src[(size_t)k];
              for (int z = 0; z < SLOT_SRCS; ++z) {
                if ((size_t)z < src.size()) {
                  const uint32_t rel = src[(size_t)z] >> 31, le = (src[(size_t)z] >> 12) & 0x7FFFFu;
                  const uint32_t i = S.slot_ij[(size_t)k][(size_t)z] >> 3, j = S.slot_ij[(size_t)k][(size_t)z] & 7u;
                  const uint32_t a = std::min(i, j), b = std::max(i, j), t = (k == len) ? 0u : a * (uint32_t)nd - (a * (a - 1)) / 2 + (b - a);
                  put(((rel * cap + (le - anchor)) * 8u) | (t << MSRC_T_SHIFT) | (j << MSRC_J_SHIFT) | (i << MSRC_I_SHIFT));
                } else put(SRC_NONE);
              }
            }
          int32_t pid;
          auto itc = S.cache.find(key);
          if (itc != S.cache.end()) pid = itc->second;
          else {
            std::lock_guard<std::mutex> lock(mu);
            auto itg = pattern_ids.find(key);
            if (itg != pattern_ids.end()) pid = itg->second;
            else {
              pid = (int32_t)out.patterns.size();
              PatternRec PR;
              PR.desc_begin = (int32_t)(out.desc[0].size() / SLOT_SRCS); PR.n_slots = len + 1;
              out.patterns.push_back(PR);
              const uint32_t* w = (const uint32_t*)key.data();
              for (size_t z = 0; z < legacy_words; ++z) {
                const uint32_t s0 = w[z];
                uint32_t s1 = s0;
                if (s0 != SRC_NONE) s1 = (s0 >= slot_bytes) ? s0 - slot_bytes : s0 + slot_bytes;
                out.desc[0].push_back(s0);
                out.desc[1].push_back(s1);
              }
              for (size_t z = legacy_words; z < key.size() / 4; ++z) {   // parity 1: the current step writes the odd slot
                const uint32_t s0 = w[z];
                uint32_t s1 = s0;
                if (s0 != SRC_NONE) {
                  const uint32_t off = s0 & MSRC_OFF_MASK;
                  s1 = (s0 & ~MSRC_OFF_MASK) | (off >= cap * 8u ? off - cap * 8u : off + cap * 8u);
                }
                out.mdesc[0].push_back(s0);
                out.mdesc[1].push_back(s1);
              }
              pattern_ids.emplace(key, pid);
            }
            S.cache.emplace(key, pid);
          }
          SR.pid = pid; SR.rec.anchor = (uint16_t)anchor;
          step_rows.push_back(SR);
        }
        // batches: up to 32 rows of one pattern (fixed rows together), rows ascending inside a batch
        std::stable_sort(step_rows.begin(), step_rows.end(), [](const StepRow& a, const StepRow& b) { return a.pid < b.pid; });
        ST.batch_begin = (int32_t)W.batches.size();
        size_t k = 0;
        while (k < step_rows.size()) {
          size_t k2 = k;
          while (k2 < step_rows.size() && step_rows[k2].pid == step_rows[k].pid && k2 - k < 32) ++k2;
          BatchRec B;
          B.row_begin = (int32_t)W.rows.size(); B.n_rows = (uint16_t)(k2 - k);
          if (step_rows[k].pid < 0) { B.desc_begin = 0; B.n_slots = 0; B.flags = BATCH_FIXED; }
          else {
            std::lock_guard<std::mutex> lock(mu);   // patterns may be reallocated by other threads
            const PatternRec& PR = out.patterns[(size_t)step_rows[k].pid];
            B.desc_begin = PR.desc_begin; B.n_slots = (uint16_t)PR.n_slots; B.flags = 0;
          }
          for (size_t q = k; q < k2; ++q) W.rows.push_back(step_rows[q].rec);
          W.batches.push_back(B);
          k = k2;
        }
        ST.n_batches = (int32_t)W.batches.size() - ST.batch_begin;
        const int32_t nrows_step = (int32_t)step_rows.size();
        int32_t cur = max_rows_step.load();
        while (nrows_step > cur && !max_rows_step.compare_exchange_weak(cur, nrows_step)) {}
        cur = max_batches_step.load();
        while (ST.n_batches > cur && !max_batches_step.compare_exchange_weak(cur, ST.n_batches)) {}
      }
      if (rk != rows_sorted.size()) throw std::runtime_error("plan: a row was not assigned to a sweep step (internal error)");
    });
    out.max_rows_step = max_rows_step.load();
    out.max_batches_step = max_batches_step.load();

    phase.lap("pass 2 (rows, patterns)");
    // ---- concatenate
    out.chain_step_ptr.assign((size_t)nchains + 1, 0);
    for (int32_t c = 0; c < nchains; ++c) {
      ChainWork& W = work[(size_t)c];
      const int32_t eb = (int32_t)out.step_elems.size(), bb = (int32_t)out.batches.size(), rb = (int32_t)out.rows.size();
      for (StepRec S : W.steps) { S.elem_begin += eb; S.batch_begin += bb; out.steps.push_back(S); }
      for (BatchRec B : W.batches) { B.row_begin += rb; out.batches.push_back(B); }
      out.step_elems.insert(out.step_elems.end(), W.elems.begin(), W.elems.end());
      out.rows.insert(out.rows.end(), W.rows.begin(), W.rows.end());
      out.chain_step_ptr[(size_t)c + 1] = (int32_t)out.steps.size();
      W = ChainWork();
    }
    out.n_elem_with_halo = (int64_t)out.step_elems.size();
    // rows: 32 per batch (padded), with the CSR offset of the row copied in, so the device needs one load level
    {
      std::vector<RowRec> padded((size_t)out.batches.size() * 32);
      RowRec zero; zero.row = 0; zero.anchor = 0; zero.aux = 0; zero.base = 0;
      for (size_t b = 0; b < out.batches.size(); ++b) {
        BatchRec& B = out.batches[b];
        for (int l = 0; l < 32; ++l) {
          RowRec R = zero;
          if (l < (int)B.n_rows) { R = out.rows[(size_t)B.row_begin + (size_t)l]; R.base = m.rowptr[(size_t)R.row]; }
          padded[b * 32 + (size_t)l] = R;
        }
        B.row_begin = (int32_t)(b * 32);
      }
      out.rows.swap(padded);
    }
    // element inputs in step order
    out.step_conn.resize((size_t)out.n_elem_with_halo * nv);
    out.step_lids.resize((size_t)out.n_elem_with_halo * nd);
    out.step_eclass.resize((size_t)out.n_elem_with_halo);
    for (int64_t k = 0; k < out.n_elem_with_halo; ++k) {
      const int32_t e = out.step_elems[(size_t)k];
      std::memcpy(&out.step_conn[(size_t)k * nv], &m.conn[(size_t)e * nv], sizeof(int32_t) * (size_t)nv);
      std::memcpy(&out.step_lids[(size_t)k * nd], &m.lids[(size_t)e * nd], sizeof(int32_t) * (size_t)nd);
      out.step_eclass[(size_t)k] = m.eclass.empty() ? 0 : m.eclass[(size_t)e];
    }
    phase.lap("concatenate + step inputs");
    // step-invariant axes per chain (extruded columns: the sweep changes one coordinate only)
    out.chain_invariant.assign((size_t)out.n_chains, 0);
    if (nv == (1 << m.dim)) {
      static const int nb[3] = {1, 3, 4};   // +xi, +eta, +zeta neighbours of vertex 0 (Shards order)
      for (int32_t c = 0; c < out.n_chains; ++c) {
        const int32_t b = out.chain_step_ptr[(size_t)c], e = out.chain_step_ptr[(size_t)c + 1];
        uint8_t inv = (uint8_t)((1 << m.dim) - 1);
        const StepRec& S0 = out.steps[(size_t)b];
        for (int32_t s = b + 1; s < e && inv; ++s) {
          const StepRec& S = out.steps[(size_t)s];
          if (S.n_elem != S0.n_elem) { inv = 0; break; }
          for (int32_t t = 0; t < S.n_elem && inv; ++t) {
            const int32_t* c0 = &out.step_conn[(size_t)(S0.elem_begin + t) * nv];
            const int32_t* c1 = &out.step_conn[(size_t)(S.elem_begin + t) * nv];
            if (out.step_eclass[(size_t)(S0.elem_begin + t)] != out.step_eclass[(size_t)(S.elem_begin + t)]) { inv = 0; break; }   // box code fills the cache
            for (int d = 0; d < m.dim; ++d) {
              const std::vector<double>& X = m.vcoord[d];
              if (std::memcmp(&X[(size_t)c0[0]], &X[(size_t)c1[0]], 8) != 0 || std::memcmp(&X[(size_t)c0[nb[d]]], &X[(size_t)c1[nb[d]]], 8) != 0) inv &= (uint8_t)~(1 << d);
            }
          }
        }
        if (e - b < 2) inv = 0;
        // bit 4 + d: in every step of the chain all elements span the same axis-d interval (a level of an extruded mesh)
        uint8_t uni = (uint8_t)((1 << m.dim) - 1);
        for (int32_t s = b; s < e && uni; ++s) {
          const StepRec& S = out.steps[(size_t)s];
          const int32_t* c0 = &out.step_conn[(size_t)S.elem_begin * nv];
          for (int32_t t = 1; t < S.n_elem && uni; ++t) {
            const int32_t* c1 = &out.step_conn[(size_t)(S.elem_begin + t) * nv];
            for (int d = 0; d < m.dim; ++d) {
              const std::vector<double>& X = m.vcoord[d];
              if (std::memcmp(&X[(size_t)c0[0]], &X[(size_t)c1[0]], 8) != 0 || std::memcmp(&X[(size_t)c0[nb[d]]], &X[(size_t)c1[nb[d]]], 8) != 0) uni &= (uint8_t)~(1 << d);
            }
          }
        }
        out.chain_invariant[(size_t)c] = (uint8_t)(inv | (uni << 4));
      }
    }
    break;
  }
}

void host_apply_chain_plan(const MeshGraph& m, const ChainPlan& cp, const double* stage, bool accumulate, double* res, double* jac) {
  const int SL = cp.stage_len;
  const int64_t slot_doubles = (int64_t)cp.cap * SL;
  std::vector<double> ring((size_t)(2 * slot_doubles), 0.0);
  for (int32_t c = 0; c < cp.n_chains; ++c) {
    for (int32_t s = cp.chain_step_ptr[(size_t)c]; s < cp.chain_step_ptr[(size_t)c + 1]; ++s) {
      const StepRec& ST = cp.steps[(size_t)s];
      const int parity = (s - cp.chain_step_ptr[(size_t)c]) & 1;
      for (int32_t l = 0; l < ST.n_elem; ++l) {
        const int32_t e = cp.step_elems[(size_t)(ST.elem_begin + l)];
        for (int t = 0; t < SL; ++t) ring[(size_t)(parity * slot_doubles + (int64_t)t * cp.cap + l)] = stage[(size_t)e * SL + t];
      }
      for (int32_t b = 0; b < ST.n_batches; ++b) {
        const BatchRec& B = cp.batches[(size_t)(ST.batch_begin + b)];
        for (int32_t lane = 0; lane < (int32_t)B.n_rows; ++lane) {
          const RowRec& R = cp.rows[(size_t)(ST.batch_begin + b) * 32 + (size_t)lane];
          const int64_t base = R.base;
          if (base != m.rowptr[(size_t)R.row]) throw std::runtime_error("plan: stale CSR offset in a row record (internal error)");
          const int32_t len = (int32_t)(m.rowptr[(size_t)R.row + 1] - base);
          if (B.flags & BATCH_FIXED) {
            if (!accumulate) {
              if (jac) for (int32_t k = 0; k < len; ++k) jac[base + k] = (k == (int32_t)R.aux) ? 1.0 : 0.0;
              if (res) res[R.row] = 0.0;
            }
            continue;
          }
          if ((int32_t)B.n_slots != len + 1) throw std::runtime_error("plan: batch pattern does not match the row length (internal error)");
          for (int32_t k = 0; k <= len; ++k) {
            double acc = 0.0;
            for (int z = 0; z < SLOT_SRCS; ++z) {
              const uint32_t src = cp.desc[parity][((size_t)B.desc_begin + (size_t)k) * SLOT_SRCS + (size_t)z];
              if (src != SRC_NONE) acc += ring[(size_t)(src / 8 + R.anchor)];
            }
            if (k == len) { if (res) { if (accumulate) res[R.row] += -acc; else res[R.row] = -acc; } }
            else if (jac) { if (accumulate) jac[base + k] += acc; else jac[base + k] = acc; }
          }
        }
      }
    }
  }
}

void host_apply_metric_plan(const MeshGraph& m, const ChainPlan& cp, const double* metric, int ng, const double* Stab, const double* Mtab,
                            double alpha_u, double alpha_t, bool accumulate, double* res, double* jac) {
  if (cp.mdesc[0].size() != cp.desc[0].size()) throw std::runtime_error("plan: no metric source words for this plan");
  const int nd = m.ndof, nt = nd * (nd + 1) / 2, SL = ng + 1 + 3 * nd;
  const int64_t cap = cp.cap, ES = 2 * cap;   // doubles between consecutive entries of one element column
  const int MD = ng, B0 = ng + 1, U0 = B0 + nd, UT0 = U0 + nd;
  std::vector<double> ring((size_t)(ES * SL), 0.0);
  for (int32_t c = 0; c < cp.n_chains; ++c) {
    for (int32_t s = cp.chain_step_ptr[(size_t)c]; s < cp.chain_step_ptr[(size_t)c + 1]; ++s) {
      const StepRec& ST = cp.steps[(size_t)s];
      const int parity = (s - cp.chain_step_ptr[(size_t)c]) & 1;
      for (int32_t l = 0; l < ST.n_elem; ++l) {
        const int32_t e = cp.step_elems[(size_t)(ST.elem_begin + l)];
        for (int t = 0; t < SL; ++t) ring[(size_t)((int64_t)t * ES + parity * cap + l)] = metric[(size_t)e * SL + t];
      }
      for (int32_t b = 0; b < ST.n_batches; ++b) {
        const BatchRec& B = cp.batches[(size_t)(ST.batch_begin + b)];
        for (int32_t lane = 0; lane < (int32_t)B.n_rows; ++lane) {
          const RowRec& R = cp.rows[(size_t)(ST.batch_begin + b) * 32 + (size_t)lane];
          const int64_t base = R.base;
          const int32_t len = (int32_t)(m.rowptr[(size_t)R.row + 1] - base);
          if (B.flags & BATCH_FIXED) {
            if (!accumulate) {
              if (jac) for (int32_t k = 0; k < len; ++k) jac[base + k] = (k == (int32_t)R.aux) ? 1.0 : 0.0;
              if (res) res[R.row] = 0.0;
            }
            continue;
          }
          double racc = 0.0;
          for (int32_t k = 0; k <= len; ++k) {
            const uint32_t* w = &cp.mdesc[parity][((size_t)B.desc_begin + (size_t)k) * SLOT_SRCS];
            if (k == len) {
              double bs = 0.0;
              for (int z = 0; z < SLOT_SRCS && w[z] != SRC_NONE; ++z)
                bs += ring[(size_t)((w[z] & MSRC_OFF_MASK) / 8 + R.anchor + (int64_t)(B0 + (int)((w[z] >> MSRC_I_SHIFT) & 7u)) * ES)];
              if (res) { const double v = bs - racc; if (accumulate) res[R.row] += v; else res[R.row] = v; }
              break;
            }
            double kp = 0.0, mp = 0.0;
            for (int z = 0; z < SLOT_SRCS && w[z] != SRC_NONE; ++z) {
              const int64_t col = (w[z] & MSRC_OFF_MASK) / 8 + R.anchor;
              const int t = (int)((w[z] >> MSRC_T_SHIFT) & 63u);
              for (int g = 0; g < ng; ++g) kp += ring[(size_t)(col + (int64_t)g * ES)] * Stab[(size_t)g * nt + t];
              mp += ring[(size_t)(col + (int64_t)MD * ES)] * Mtab[t];
            }
            const int64_t col0 = (w[0] & MSRC_OFF_MASK) / 8 + R.anchor;
            const int j0 = (int)((w[0] >> MSRC_J_SHIFT) & 7u);
            racc += kp * ring[(size_t)(col0 + (int64_t)(U0 + j0) * ES)] + mp * ring[(size_t)(col0 + (int64_t)(UT0 + j0) * ES)];
            if (jac) { const double v = alpha_u * kp + alpha_t * mp; if (accumulate) jac[base + k] += v; else jac[base + k] = v; }
          }
        }
      }
    }
  }
}

}  // namespace mrhyde_b200
