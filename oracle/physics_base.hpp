// ORACLE (test infrastructure, never shipped, never on the product path).
//
// Plugin interface restated from PhysicsBase<EvalT> (src/physics/physicsBase.hpp:29-206):
// the virtuals defineFunctions / volumeResidual / boundaryResidual / setWorkset, the
// per-module variable and basis-type lists, and the flat string settings standing in for
// the module's Teuchos::ParameterList.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "workset.hpp"

namespace oracle {

struct Settings {  // flattened YAML: "Sublist/Sublist/key" -> value
  std::map<std::string, std::string> kv;
  bool has(const std::string& k) const { return kv.count(k) > 0; }
  std::string get(const std::string& k, const std::string& def) const { auto it = kv.find(k); return it == kv.end() ? def : it->second; }
  double getd(const std::string& k, double def) const { auto it = kv.find(k); return it == kv.end() ? def : std::stod(it->second); }
  int geti(const std::string& k, int def) const { auto it = kv.find(k); return it == kv.end() ? def : std::stoi(it->second); }
  bool getb(const std::string& k, bool def) const {
    auto it = kv.find(k);
    if (it == kv.end()) return def;
    return it->second == "true" || it->second == "True" || it->second == "1";
  }
  // all (suffix, value) pairs below a prefix "A/B/"
  std::vector<std::pair<std::string, std::string>> sub(const std::string& prefix) const {
    std::vector<std::pair<std::string, std::string>> out;
    for (auto& p : kv) if (p.first.compare(0, prefix.size(), prefix) == 0) out.push_back({p.first.substr(prefix.size()), p.second});
    return out;
  }
};

template <class EvalT>
class PhysicsBase {
 public:
  std::string label;
  std::vector<std::string> myvars, mybasistypes;
  Workset<EvalT>* wkset = nullptr;
  FunctionManager<EvalT>* functionManager = nullptr;
  virtual ~PhysicsBase() {}
  virtual void defineFunctions(const Settings& fs, FunctionManager<EvalT>* fm) = 0;
  virtual void volumeResidual() = 0;
  virtual void boundaryResidual() {}
  virtual void setWorkset(Workset<EvalT>* w) { wkset = w; }
  int findVar(const std::string& name) const {
    for (size_t i = 0; i < wkset->vars.size(); ++i) if (wkset->vars[i].name == name) return (int)i;
    return -1;
  }
};

}  // namespace oracle
