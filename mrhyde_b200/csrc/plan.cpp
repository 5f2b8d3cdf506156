// Host-side plan construction (see plan.hpp).
#include "plan.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
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

inline uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3ff;
  v = (v | (v << 16)) & 0x030000FF;
  v = (v | (v << 8)) & 0x0300F00F;
  v = (v | (v << 4)) & 0x030C30C3;
  v = (v | (v << 2)) & 0x09249249;
  return v;
}

template <class F>
void parallel_for(int64_t n, F&& f) {
  int nt = (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 16) nt = 16;
  if (n < 4 * nt) nt = 1;
  std::atomic<int64_t> next(0);
  auto worker = [&](int tid) {
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n) break;
      f(i, tid);
    }
  };
  if (nt == 1) { worker(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) th.emplace_back(worker, t);
  for (auto& t : th) t.join();
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

void MeshGraph::classify_affine() {
  affine.assign((size_t)nelem, 0);
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
    affine[(size_t)e] = (worst <= tol * emin) ? 1 : 0;
  }
}

void build_patch_plan(const MeshGraph& m, const std::vector<uint16_t>& kmap, const std::vector<uint16_t>& rmap,
                      int stage_len, int chunk_target, size_t smem_budget_bytes, PatchPlan& out) {
  const int64_t ne = m.nelem, nr = m.nrows;
  const int nd = m.ndof, nv = m.nverts;
  if (ne <= 0 || nr <= 0) throw std::runtime_error("plan: empty mesh or graph");

  // ---- Morton order of element centroids
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int d = 0; d < m.dim; ++d)
    for (int64_t v = 0; v < m.nvert; ++v) { lo[d] = std::min(lo[d], m.vcoord[d][v]); hi[d] = std::max(hi[d], m.vcoord[d][v]); }
  std::vector<uint64_t> keyed((size_t)ne);
  for (int64_t e = 0; e < ne; ++e) {
    uint32_t key = 0;
    for (int d = 0; d < m.dim; ++d) {
      double c = 0.0;
      for (int n = 0; n < nv; ++n) c += m.vcoord[d][m.conn[(size_t)e * nv + n]];
      c /= nv;
      const double t = (hi[d] > lo[d]) ? (c - lo[d]) / (hi[d] - lo[d]) : 0.0;
      uint32_t q = (uint32_t)std::min(1023.0, std::max(0.0, t * 1024.0));
      key |= spread10(q) << d;
    }
    keyed[(size_t)e] = ((uint64_t)key << 32) | (uint64_t)e;
  }
  std::sort(keyed.begin(), keyed.end());
  std::vector<int32_t> pos((size_t)ne);
  for (int64_t k = 0; k < ne; ++k) pos[(size_t)(keyed[(size_t)k] & 0xffffffffu)] = (int32_t)k;

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

  int chunk = std::max(1, chunk_target);
  for (;;) {  // shrink the chunk until every patch fits the shared-memory budget and 16-bit staging indices
    out = PatchPlan();
    out.chunk = chunk;
    const int32_t npatch = (int32_t)((ne + chunk - 1) / chunk);
    // owner patch of each row = largest chunk among its elements
    std::vector<int32_t> owner((size_t)nr, -1);
    for (int64_t r = 0; r < nr; ++r)
      for (int64_t p = r2e_ptr[(size_t)r]; p < r2e_ptr[(size_t)r + 1]; ++p)
        owner[(size_t)r] = std::max(owner[(size_t)r], pos[(size_t)r2e_elem[(size_t)p]] / chunk);
    out.patch_row_ptr.assign((size_t)npatch + 1, 0);
    for (int64_t r = 0; r < nr; ++r) {
      if (owner[(size_t)r] >= 0) ++out.patch_row_ptr[(size_t)owner[(size_t)r] + 1];
      else out.orphan_rows.push_back((int32_t)r);
    }
    for (int32_t p = 0; p < npatch; ++p) out.patch_row_ptr[(size_t)p + 1] += out.patch_row_ptr[(size_t)p];
    out.patch_rows.assign((size_t)out.patch_row_ptr[(size_t)npatch], 0);
    {
      std::vector<int32_t> fill(out.patch_row_ptr.begin(), out.patch_row_ptr.end() - 1);
      for (int64_t r = 0; r < nr; ++r) if (owner[(size_t)r] >= 0) out.patch_rows[(size_t)fill[(size_t)owner[(size_t)r]]++] = (int32_t)r;
    }

    // per-patch element lists (halo included), ordered by Morton position
    std::vector<std::vector<int32_t>> pelems((size_t)npatch);
    int nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 16) nthreads = 16;
    std::vector<std::vector<int32_t>> stamp((size_t)nthreads);
    std::atomic<int> too_big(0);
    parallel_for(npatch, [&](int64_t p, int tid) {
      auto& st = stamp[(size_t)tid];
      if (st.empty()) st.assign((size_t)ne, -1);
      auto& list = pelems[(size_t)p];
      for (int32_t k = out.patch_row_ptr[(size_t)p]; k < out.patch_row_ptr[(size_t)p + 1]; ++k) {
        const int32_t r = out.patch_rows[(size_t)k];
        for (int64_t q = r2e_ptr[(size_t)r]; q < r2e_ptr[(size_t)r + 1]; ++q) {
          const int32_t e = r2e_elem[(size_t)q];
          if (st[(size_t)e] != (int32_t)p) { st[(size_t)e] = (int32_t)p; list.push_back(e); }
        }
      }
      std::sort(list.begin(), list.end(), [&](int32_t a, int32_t b) { return pos[(size_t)a] < pos[(size_t)b]; });
      if ((size_t)list.size() * stage_len * sizeof(double) > smem_budget_bytes || (size_t)list.size() * stage_len > 65535u) too_big = 1;
    });
    if (too_big.load()) {
      if (chunk == 1) throw std::runtime_error("plan: a single element patch exceeds the shared-memory budget");
      chunk = std::max(1, chunk / 2);
      continue;
    }
    out.n_patches = npatch;
    out.patch_elem_ptr.assign((size_t)npatch + 1, 0);
    for (int32_t p = 0; p < npatch; ++p) out.patch_elem_ptr[(size_t)p + 1] = out.patch_elem_ptr[(size_t)p] + (int32_t)pelems[(size_t)p].size();
    out.patch_elems.resize((size_t)out.patch_elem_ptr[(size_t)npatch]);
    for (int32_t p = 0; p < npatch; ++p) {
      std::copy(pelems[(size_t)p].begin(), pelems[(size_t)p].end(), out.patch_elems.begin() + out.patch_elem_ptr[(size_t)p]);
      out.max_pe = std::max(out.max_pe, (int32_t)pelems[(size_t)p].size());
      out.max_rows = std::max(out.max_rows, out.patch_row_ptr[(size_t)p + 1] - out.patch_row_ptr[(size_t)p]);
    }
    out.n_elem_with_halo = out.patch_elem_ptr[(size_t)npatch];

    // ---- scatter programs, de-duplicated into templates
    out.patch_tmpl.assign((size_t)npatch, -1);
    std::unordered_map<uint64_t, std::vector<int32_t>> by_hash;
    std::mutex mu;
    struct Scratch {
      std::vector<int32_t> local;      // global element -> local index (stamped)
      std::vector<int32_t> local_tag;
      std::vector<uint16_t> slot_row, slot_k, csrc;
      std::vector<uint32_t> cptr;
      std::vector<uint32_t> ckey;      // per-row scratch: (slot << 16) | src
    };
    std::vector<Scratch> scratch((size_t)nthreads);
    parallel_for(npatch, [&](int64_t p, int tid) {
      Scratch& S = scratch[(size_t)tid];
      if (S.local.empty()) { S.local.assign((size_t)ne, 0); S.local_tag.assign((size_t)ne, -1); }
      const auto& list = pelems[(size_t)p];
      const int32_t n_pe = (int32_t)list.size();
      for (int32_t l = 0; l < n_pe; ++l) { S.local[(size_t)list[(size_t)l]] = l; S.local_tag[(size_t)list[(size_t)l]] = (int32_t)p; }
      S.slot_row.clear(); S.slot_k.clear(); S.csrc.clear(); S.cptr.clear();
      const int32_t r0 = out.patch_row_ptr[(size_t)p], r1 = out.patch_row_ptr[(size_t)p + 1];
      for (int32_t lr = 0; lr < r1 - r0; ++lr) {
        const int32_t r = out.patch_rows[(size_t)(r0 + lr)];
        const int64_t rs = m.rowptr[(size_t)r], re = m.rowptr[(size_t)r + 1];
        const int32_t len = (int32_t)(re - rs);
        const int32_t* cols = &m.colind[(size_t)rs];
        // contributions of this row, keyed by slot; adjacency is element-ascending so a stable sort keeps that order
        S.ckey.clear();
        for (int64_t q = r2e_ptr[(size_t)r]; q < r2e_ptr[(size_t)r + 1]; ++q) {
          const int32_t e = r2e_elem[(size_t)q];
          const int i = r2e_dof[(size_t)q];
          const int32_t le = S.local[(size_t)e];
          for (int j = 0; j < nd; ++j) {
            const int32_t c = m.lids[(size_t)e * nd + j];
            const int32_t* it = std::lower_bound(cols, cols + len, c);
            if (it == cols + len || *it != c) throw std::runtime_error("plan: graph is missing an element coupling (row " + std::to_string(r) + ", col " + std::to_string(c) + ")");
            const uint32_t k = (uint32_t)(it - cols);
            S.ckey.push_back((k << 16) | (uint32_t)((uint32_t)kmap[(size_t)i * nd + j] * n_pe + le));
          }
          S.ckey.push_back(((uint32_t)len << 16) | (uint32_t)((uint32_t)rmap[(size_t)i] * n_pe + le));
        }
        std::stable_sort(S.ckey.begin(), S.ckey.end(), [](uint32_t a, uint32_t b) { return (a >> 16) < (b >> 16); });
        size_t c = 0;
        for (int32_t k = 0; k <= len; ++k) {
          S.slot_row.push_back((uint16_t)lr);
          S.slot_k.push_back(k == len ? SLOT_RES : (uint16_t)k);
          S.cptr.push_back((uint32_t)S.csrc.size());
          while (c < S.ckey.size() && (int32_t)(S.ckey[c] >> 16) == k) { S.csrc.push_back((uint16_t)(S.ckey[c] & 0xffff)); ++c; }
        }
      }
      S.cptr.push_back((uint32_t)S.csrc.size());
      // hash + de-duplicate
      uint64_t h = mix64((uint64_t)n_pe * 1315423911u + (uint64_t)(r1 - r0));
      auto feed = [&](const void* d, size_t nbytes) {
        const uint8_t* b = (const uint8_t*)d;
        size_t i = 0;
        for (; i + 8 <= nbytes; i += 8) { uint64_t w; std::memcpy(&w, b + i, 8); h = mix64(h ^ w) + 0x9e3779b97f4a7c15ULL; }
        uint64_t w = 0; if (i < nbytes) { std::memcpy(&w, b + i, nbytes - i); h = mix64(h ^ w) + 0x51ed270b; }
      };
      feed(S.slot_row.data(), S.slot_row.size() * 2); feed(S.slot_k.data(), S.slot_k.size() * 2);
      feed(S.cptr.data(), S.cptr.size() * 4); feed(S.csrc.data(), S.csrc.size() * 2);
      std::lock_guard<std::mutex> lock(mu);
      int32_t found = -1;
      for (int32_t t : by_hash[h]) {
        const TemplateHeader& T = out.tmpl[(size_t)t];
        if (T.n_pe != n_pe || T.n_rows != r1 - r0 || T.n_slots != (int32_t)S.slot_row.size()) continue;
        const uint32_t nc = out.cptr[(size_t)T.off_cptr + T.n_slots];
        if (nc != S.csrc.size()) continue;
        if (std::memcmp(&out.slot_row[(size_t)T.off_slot], S.slot_row.data(), S.slot_row.size() * 2)) continue;
        if (std::memcmp(&out.slot_k[(size_t)T.off_slot], S.slot_k.data(), S.slot_k.size() * 2)) continue;
        if (std::memcmp(&out.cptr[(size_t)T.off_cptr], S.cptr.data(), S.cptr.size() * 4)) continue;
        if (std::memcmp(&out.csrc[(size_t)T.off_csrc], S.csrc.data(), S.csrc.size() * 2)) continue;
        found = t; break;
      }
      if (found < 0) {
        TemplateHeader T;
        T.n_pe = n_pe; T.n_rows = r1 - r0; T.n_slots = (int32_t)S.slot_row.size(); T.pad = 0;
        T.off_slot = (int64_t)out.slot_row.size(); T.off_cptr = (int64_t)out.cptr.size(); T.off_csrc = (int64_t)out.csrc.size();
        out.slot_row.insert(out.slot_row.end(), S.slot_row.begin(), S.slot_row.end());
        out.slot_k.insert(out.slot_k.end(), S.slot_k.begin(), S.slot_k.end());
        out.cptr.insert(out.cptr.end(), S.cptr.begin(), S.cptr.end());
        out.csrc.insert(out.csrc.end(), S.csrc.begin(), S.csrc.end());
        found = (int32_t)out.tmpl.size();
        out.tmpl.push_back(T);
        by_hash[h].push_back(found);
      }
      out.patch_tmpl[(size_t)p] = found;
      out.max_slots = std::max(out.max_slots, (int32_t)S.slot_row.size());
    });
    break;
  }
}

}  // namespace mrhyde_b200
