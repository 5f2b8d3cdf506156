// Plan-time analysis of the general path's deterministic pull (see general.hpp).
//
// The reference adds element matrices into the CSR matrix entry by entry with a linear search per entry and
// atomics when threaded (assemblyManager_scatter.hpp:191-278).  Here every (instance, local row) pair is listed
// under its global row in ascending instance order -- volume elements in element order, then boundary sides in
// boundary-group order, which is the order the serial reference adds them (assemblyManager_jacres.hpp:336-603) --
// and the position of every local column inside that CSR row is tabulated once (uint16).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <unordered_map>

#include "general.hpp"

namespace mrhyde_b200 {

void gen_build_pull(const MeshGraph& m, int N, const std::vector<GenSideFamily>& sides, int64_t batch_elems, int64_t scratch_budget_bytes, GeneralPlanHost& out) {
  const int64_t ne = m.nelem;
  int64_t n_inst = ne;
  for (auto& s : sides) if (s.active) n_inst += (int64_t)s.items.size();
  if (n_inst * N > 0x7fffffffLL) throw std::runtime_error("general plan: element instances x dofs exceed the 32-bit contribution index");
  const int64_t inst_bytes = (int64_t)N * (N + 1) * 8;   // element matrix + vector
  if (batch_elems <= 0 && scratch_budget_bytes > 0 && n_inst * inst_bytes > scratch_budget_bytes) {
    // a ring of two batches (what an element numbering that sweeps the mesh needs) plus the side instances must fit the budget
    batch_elems = std::max<int64_t>(1024, (scratch_budget_bytes / inst_bytes - (n_inst - ne)) / 2);
  }
  out.n_elem = ne; out.n_inst = n_inst; out.n_rows = m.nrows; out.n_owned = m.nowned;
  // instance -> element
  std::vector<int32_t> inst_elem((size_t)n_inst);
  for (int64_t e = 0; e < ne; ++e) inst_elem[(size_t)e] = (int32_t)e;
  for (auto& s : sides) {
    if (!s.active) continue;
    for (size_t k = 0; k < s.items.size(); ++k) inst_elem[(size_t)(s.inst_base + (int64_t)k)] = s.items[k];
  }
  // rows -> (instance, local row) lists, ascending instance; completion instance of every row
  std::vector<int64_t> cnt((size_t)m.nrows + 1, 0);
  for (int64_t t = 0; t < n_inst; ++t) {
    const int32_t* l = &m.lids[(size_t)inst_elem[(size_t)t] * N];
    for (int i = 0; i < N; ++i) ++cnt[(size_t)l[i] + 1];
  }
  for (int64_t r = 0; r < m.nrows; ++r) cnt[(size_t)r + 1] += cnt[(size_t)r];
  std::vector<int32_t> contrib_by_row((size_t)cnt[(size_t)m.nrows]);
  std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
  std::vector<int64_t> last_inst((size_t)m.nrows, -1);
  for (int64_t t = 0; t < n_inst; ++t) {
    const int32_t* l = &m.lids[(size_t)inst_elem[(size_t)t] * N];
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < i; ++j)
        if (l[j] == l[i]) throw std::runtime_error("general plan: an element lists the same dof twice (degenerate periodic mesh?)");
      contrib_by_row[(size_t)fill[(size_t)l[i]]++] = (int32_t)(t * N + i);
      last_inst[(size_t)l[i]] = t;
    }
  }
  // volume batches over elements
  out.batches.clear();
  if (batch_elems <= 0 || batch_elems > ne) batch_elems = std::max<int64_t>(ne, 1);
  for (int64_t e0 = 0; e0 < ne; e0 += batch_elems) {
    GenBatch b;
    b.elem_begin = e0; b.elem_end = std::min(ne, e0 + batch_elems);
    out.batches.push_back(b);
  }
  if (out.batches.empty()) out.batches.push_back(GenBatch());
  const int nb = (int)out.batches.size();
  auto batch_of = [&](int64_t inst) -> int {
    if (inst < 0) return 0;
    if (inst >= ne) return nb - 1;
    return (int)std::min<int64_t>(inst / batch_elems, nb - 1);
  };
  // rows sorted by completion batch (stable: ascending row inside a batch)
  std::vector<int64_t> bcount((size_t)nb + 1, 0);
  for (int64_t r = 0; r < m.nrows; ++r) ++bcount[(size_t)batch_of(last_inst[(size_t)r]) + 1];
  for (int b = 0; b < nb; ++b) bcount[(size_t)b + 1] += bcount[(size_t)b];
  for (int b = 0; b < nb; ++b) { out.batches[(size_t)b].row_begin = bcount[(size_t)b]; out.batches[(size_t)b].row_end = bcount[(size_t)b + 1]; }
  out.row_order.assign((size_t)m.nrows, 0);
  {
    std::vector<int64_t> bf(bcount.begin(), bcount.end() - 1);
    for (int64_t r = 0; r < m.nrows; ++r) out.row_order[(size_t)bf[(size_t)batch_of(last_inst[(size_t)r])]++] = (int32_t)r;
  }
  out.contrib_ptr.assign((size_t)m.nrows + 1, 0);
  out.contrib.resize(contrib_by_row.size());
  int64_t w = 0;
  int32_t maxlen = 0;
  for (int64_t k = 0; k < m.nrows; ++k) {
    const int32_t r = out.row_order[(size_t)k];
    out.contrib_ptr[(size_t)k] = w;
    for (int64_t p = cnt[(size_t)r]; p < cnt[(size_t)r + 1]; ++p) out.contrib[(size_t)w++] = contrib_by_row[(size_t)p];
    maxlen = std::max<int32_t>(maxlen, (int32_t)(m.rowptr[(size_t)r + 1] - m.rowptr[(size_t)r]));
  }
  out.contrib_ptr[(size_t)m.nrows] = w;
  out.max_row_len = maxlen;
  if (maxlen > 0xffff) throw std::runtime_error("general plan: CSR rows longer than 65535 entries");
  // ---- scratch ring: volume instance t may be overwritten by instance t + cap only after the batch that pulls the last row it
  // feeds, i.e. batch_of(t + cap) > need(t) for every t.  cap = the smallest multiple of the batch size with that property.
  {
    std::vector<int32_t> need((size_t)ne, 0);   // last batch that reads instance t
    for (int64_t t = 0; t < ne; ++t) {
      const int32_t* l = &m.lids[(size_t)t * N];
      int nb_t = 0;
      for (int i = 0; i < N; ++i) nb_t = std::max(nb_t, batch_of(last_inst[(size_t)l[i]]));
      need[(size_t)t] = nb_t;
    }
    int span = 1;   // batches an instance stays live, including its own
    for (int64_t t = 0; t < ne; ++t) span = std::max(span, need[(size_t)t] - batch_of(t) + 1);
    out.scratch_cap = std::min<int64_t>(std::max<int64_t>(ne, 1), (int64_t)span * batch_elems);
    if (out.scratch_cap < ne && (out.scratch_cap + (n_inst - ne)) * N > 0x7fffffffLL) throw std::runtime_error("general plan: scratch ring x dofs exceed the 32-bit contribution index");
  }
  // ---- column positions, de-duplicated: instances whose N x N position tables coincide share one.  Every worker keeps the
  // distinct tables of its instance range; the ranges are merged afterwards (first occurrence in instance order gets the lower id).
  {
    const size_t NN = (size_t)N * N;
    const unsigned nth = (n_inst * (int64_t)NN < (1 << 22)) ? 1u : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    struct Local { std::vector<uint16_t> tab; std::vector<int32_t> id; std::string err; };
    std::vector<Local> loc(nth);
    const int64_t chunk = (n_inst + nth - 1) / nth;
    auto work = [&](unsigned k) {
      Local& L = loc[k];
      const int64_t t0 = std::min<int64_t>(n_inst, (int64_t)k * chunk), t1 = std::min<int64_t>(n_inst, (int64_t)(k + 1) * chunk);
      std::unordered_map<std::string, int32_t> ids;
      std::vector<uint16_t> cur(NN);
      L.id.reserve((size_t)(t1 - t0));
      for (int64_t t = t0; t < t1; ++t) {
        const int32_t* l = &m.lids[(size_t)inst_elem[(size_t)t] * N];
        for (int i = 0; i < N; ++i) {
          const int64_t rs = m.rowptr[(size_t)l[i]], re = m.rowptr[(size_t)l[i] + 1];
          const int32_t* cb = &m.colind[(size_t)rs];
          const int32_t* ce = cb + (re - rs);
          const bool sorted = std::is_sorted(cb, ce);
          for (int c = 0; c < N; ++c) {
            const int32_t* f = sorted ? std::lower_bound(cb, ce, l[c]) : std::find(cb, ce, l[c]);
            if (f == ce || *f != l[c]) { L.err = "general plan: the CSR graph lacks an entry for a column of one of the row's elements"; return; }
            cur[(size_t)i * N + c] = (uint16_t)(f - cb);
          }
        }
        const std::string key((const char*)cur.data(), NN * sizeof(uint16_t));
        auto it = ids.find(key);
        if (it == ids.end()) {
          it = ids.emplace(key, (int32_t)ids.size()).first;
          L.tab.insert(L.tab.end(), cur.begin(), cur.end());
        }
        L.id.push_back(it->second);
      }
    };
    if (nth == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (unsigned k = 0; k < nth; ++k) th.emplace_back(work, k);
      for (auto& t : th) t.join();
    }
    for (auto& L : loc) if (!L.err.empty()) throw std::runtime_error(L.err);
    out.pos_tab.clear();
    out.pos_id.assign((size_t)n_inst, 0);
    std::unordered_map<std::string, int32_t> ids;
    for (unsigned k = 0; k < nth; ++k) {
      Local& L = loc[k];
      const size_t nloc = L.tab.size() / NN;
      std::vector<int32_t> remap(nloc, 0);
      for (size_t q = 0; q < nloc; ++q) {
        const std::string key((const char*)&L.tab[q * NN], NN * sizeof(uint16_t));
        auto it = ids.find(key);
        if (it == ids.end()) {
          it = ids.emplace(key, (int32_t)ids.size()).first;
          out.pos_tab.insert(out.pos_tab.end(), L.tab.begin() + (ptrdiff_t)(q * NN), L.tab.begin() + (ptrdiff_t)((q + 1) * NN));
        }
        remap[q] = it->second;
      }
      const int64_t t0 = std::min<int64_t>(n_inst, (int64_t)k * chunk);
      for (size_t j = 0; j < L.id.size(); ++j) out.pos_id[(size_t)t0 + j] = remap[(size_t)L.id[j]];
      L = Local();
    }
    if ((int64_t)(out.pos_tab.size() / NN) * N > 0x7fffffffLL) throw std::runtime_error("general plan: too many distinct column-position tables");
  }
}

void gen_pull_mass_host(const GeneralPlanHost& H, const MeshGraph& m, const double* elem_jac, bool accumulate, bool lump, double* mass, double* diag) {
  const int N = H.info.N;
  std::vector<double> buf((size_t)std::max(1, H.max_row_len));
  for (int64_t k = 0; k < H.n_rows; ++k) {
    const int32_t r = H.row_order[(size_t)k];
    const int64_t rs = m.rowptr[(size_t)r];
    const int len = (int)(m.rowptr[(size_t)r + 1] - rs);
    for (int t = 0; t < len; ++t) buf[(size_t)t] = 0.0;
    double lumped = 0.0;
    int dpos = -1;
    for (int t = 0; t < len; ++t) if (m.colind[(size_t)(rs + t)] == r) dpos = t;
    for (int64_t p = H.contrib_ptr[(size_t)k]; p < H.contrib_ptr[(size_t)k + 1]; ++p) {
      const int64_t ci = H.contrib[(size_t)p];
      if (ci / N >= H.n_elem) continue;   // boundary instances carry no mass
      const uint16_t* ps = &H.pos_tab[((size_t)H.pos_id[(size_t)(ci / N)] * N + (size_t)(ci % N)) * N];
      for (int c = 0; c < N; ++c) { buf[ps[c]] += elem_jac[ci * N + c]; lumped += std::fabs(elem_jac[ci * N + c]); }
    }
    if (mass) for (int t = 0; t < len; ++t) mass[rs + t] = (accumulate ? mass[rs + t] : 0.0) + buf[(size_t)t];
    if (diag) diag[r] = (accumulate ? diag[r] : 0.0) + (lump ? lumped : (dpos >= 0 ? buf[(size_t)dpos] : 0.0));
  }
}

void gen_pull_apply_host(const GeneralPlanHost& H, const double* elem_res, bool accumulate, double* y) {
  const int N = H.info.N;
  for (int64_t k = 0; k < H.n_rows; ++k) {
    const int32_t r = H.row_order[(size_t)k];
    double s = 0.0;
    for (int64_t p = H.contrib_ptr[(size_t)k]; p < H.contrib_ptr[(size_t)k + 1]; ++p) {
      const int64_t ci = H.contrib[(size_t)p];
      if (ci / N < H.n_elem) s += elem_res[ci];
    }
    y[r] = (accumulate ? y[r] : 0.0) + s;
  }
}

void gen_pull_host(const GeneralPlanHost& H, const MeshGraph& m, const double* elem_jac, const double* elem_res, bool accumulate,
                   double* res, double* jac) {
  const int N = H.info.N;
  std::vector<double> buf((size_t)std::max(1, H.max_row_len));
  for (int64_t k = 0; k < H.n_rows; ++k) {
    const int32_t r = H.row_order[(size_t)k];
    const int64_t rs = m.rowptr[(size_t)r];
    const int len = (int)(m.rowptr[(size_t)r + 1] - rs);
    if (m.fixed[(size_t)r]) {
      if (!accumulate) {
        if (res) res[r] = 0.0;
        if (jac) for (int t = 0; t < len; ++t) jac[rs + t] = (m.colind[(size_t)(rs + t)] == r && r < m.nowned) ? 1.0 : 0.0;
      }
      continue;
    }
    for (int t = 0; t < len; ++t) buf[(size_t)t] = 0.0;
    double rsum = 0.0;
    for (int64_t p = H.contrib_ptr[(size_t)k]; p < H.contrib_ptr[(size_t)k + 1]; ++p) {
      const int64_t ci = H.contrib[(size_t)p];
      const uint16_t* ps = &H.pos_tab[((size_t)H.pos_id[(size_t)(ci / N)] * N + (size_t)(ci % N)) * N];
      if (elem_jac && jac) {
        if (H.lump_mass) {   // every entry of the element row goes to the diagonal
          int dpos = -1;
          for (int t = 0; t < len; ++t) if (m.colind[(size_t)(rs + t)] == r) dpos = t;
          if (dpos >= 0) for (int c = 0; c < N; ++c) buf[(size_t)dpos] += elem_jac[ci * N + c];
        } else {
          for (int c = 0; c < N; ++c) buf[ps[c]] += elem_jac[ci * N + c];
        }
      }
      if (elem_res) rsum += elem_res[ci];
    }
    if (jac) for (int t = 0; t < len; ++t) jac[rs + t] = (accumulate ? jac[rs + t] : 0.0) + buf[(size_t)t];
    if (res) res[r] = (accumulate ? res[r] : 0.0) + (-rsum);
  }
}

}  // namespace mrhyde_b200
