// parameter_input.cpp — input-deck parser (see pb2/parameter_input.hpp).
#include "pb2/parameter_input.hpp"

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace parthenon {
namespace {
std::string Trim(const std::string &s) {
  const auto b = s.find_first_not_of(" \t\r\n");
  if (b == std::string::npos) return "";
  const auto e = s.find_last_not_of(" \t\r\n");
  return s.substr(b, e - b + 1);
}
std::string Lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); });
  return s;
}
} // namespace

void ParameterInput::LoadFromString(const std::string &text) {
  std::istringstream in(text);
  std::string line, block;
  while (std::getline(in, line)) {
    const auto hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    line = Trim(line);
    if (line.empty()) continue;
    if (line.front() == '<') {
      const auto close = line.find('>');
      PARTHENON_REQUIRE(close != std::string::npos, "input deck: unterminated block name");
      block = Trim(line.substr(1, close - 1));
      if (!blocks_.count(block)) {
        blocks_[block];
        order_.push_back(block);
      }
      continue;
    }
    const auto eq = line.find('=');
    PARTHENON_REQUIRE(eq != std::string::npos, "input deck: expected key = value: " + line);
    PARTHENON_REQUIRE(!block.empty(), "input deck: parameter before any <block>");
    blocks_[block][Trim(line.substr(0, eq))] = Trim(line.substr(eq + 1));
  }
}

void ParameterInput::LoadFromFile(const std::string &path) {
  std::ifstream f(path);
  PARTHENON_REQUIRE(f.good(), "cannot open input file " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  LoadFromString(ss.str());
}

void ParameterInput::ModifyFromString(const std::string &a) {
  const auto eq = a.find('=');
  const auto slash = a.rfind('/', eq);
  PARTHENON_REQUIRE(eq != std::string::npos && slash != std::string::npos,
                    "command-line override must be block/key=value: " + a);
  const std::string block = Trim(a.substr(0, slash));
  if (!blocks_.count(block)) {
    blocks_[block];
    order_.push_back(block);
  }
  blocks_[block][Trim(a.substr(slash + 1, eq - slash - 1))] = Trim(a.substr(eq + 1));
}

void ParameterInput::ModifyFromCmdline(int argc, char *argv[]) {
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a.empty() || a[0] == '-') {
      if (a == "-i" || a == "-r" || a == "-d" || a == "-t") ++i; // option with a value
      continue;
    }
    if (a.find('=') != std::string::npos) ModifyFromString(a);
  }
}

const std::string *ParameterInput::Find(const std::string &block,
                                        const std::string &key) const {
  auto b = blocks_.find(block);
  if (b == blocks_.end()) return nullptr;
  auto k = b->second.find(key);
  return k == b->second.end() ? nullptr : &k->second;
}

const std::string &ParameterInput::Require(const std::string &block,
                                           const std::string &key) const {
  const std::string *v = Find(block, key);
  PARTHENON_REQUIRE(v != nullptr, "Parameter name '" + key + "' not found in block '" + block + "'");
  return *v;
}

bool ParameterInput::DoesParameterExist(const std::string &block,
                                        const std::string &key) const {
  return Find(block, key) != nullptr;
}

int ParameterInput::GetInteger(const std::string &block, const std::string &key) const {
  return static_cast<int>(std::strtol(Require(block, key).c_str(), nullptr, 10));
}
Real ParameterInput::GetReal(const std::string &block, const std::string &key) const {
  return std::strtod(Require(block, key).c_str(), nullptr);
}
bool ParameterInput::GetBoolean(const std::string &block, const std::string &key) const {
  const std::string v = Lower(Require(block, key));
  if (v == "true" || v == "1" || v == "yes" || v == "on") return true;
  if (v == "false" || v == "0" || v == "no" || v == "off") return false;
  return std::strtol(v.c_str(), nullptr, 10) != 0;
}
std::string ParameterInput::GetString(const std::string &block,
                                      const std::string &key) const {
  return Require(block, key);
}

int ParameterInput::GetOrAddInteger(const std::string &block, const std::string &key, int def) {
  if (Find(block, key)) return GetInteger(block, key);
  SetInteger(block, key, def);
  return def;
}
Real ParameterInput::GetOrAddReal(const std::string &block, const std::string &key, Real def) {
  if (Find(block, key)) return GetReal(block, key);
  SetReal(block, key, def);
  return def;
}
bool ParameterInput::GetOrAddBoolean(const std::string &block, const std::string &key,
                                     bool def) {
  if (Find(block, key)) return GetBoolean(block, key);
  SetBoolean(block, key, def);
  return def;
}
std::string ParameterInput::GetOrAddString(const std::string &block, const std::string &key,
                                           const std::string &def) {
  if (Find(block, key)) return GetString(block, key);
  SetString(block, key, def);
  return def;
}
std::string ParameterInput::GetOrAddString(const std::string &block, const std::string &key,
                                           const std::string &def,
                                           const std::vector<std::string> &allowed) {
  const std::string v = GetOrAddString(block, key, def);
  PARTHENON_REQUIRE(std::find(allowed.begin(), allowed.end(), v) != allowed.end(),
                    "Parameter '" + key + "' in block '" + block + "' has invalid value '" + v +
                        "'");
  return v;
}

void ParameterInput::SetString(const std::string &block, const std::string &key,
                               const std::string &v) {
  if (!blocks_.count(block)) order_.push_back(block);
  blocks_[block][key] = v;
}
void ParameterInput::SetInteger(const std::string &block, const std::string &key, int v) {
  SetString(block, key, std::to_string(v));
}
void ParameterInput::SetReal(const std::string &block, const std::string &key, Real v) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%.17g", v);
  SetString(block, key, buf);
}
void ParameterInput::SetBoolean(const std::string &block, const std::string &key, bool v) {
  SetString(block, key, v ? "true" : "false");
}

template <>
std::vector<std::string> ParameterInput::GetVector<std::string>(const std::string &block,
                                                                const std::string &key) const {
  std::vector<std::string> out;
  std::istringstream ss(Require(block, key));
  std::string item;
  while (std::getline(ss, item, ',')) out.push_back(Trim(item));
  return out;
}
template <>
std::vector<int> ParameterInput::GetVector<int>(const std::string &block,
                                                const std::string &key) const {
  std::vector<int> out;
  for (auto &s : GetVector<std::string>(block, key))
    out.push_back(static_cast<int>(std::strtol(s.c_str(), nullptr, 10)));
  return out;
}
template <>
std::vector<Real> ParameterInput::GetVector<Real>(const std::string &block,
                                                  const std::string &key) const {
  std::vector<Real> out;
  for (auto &s : GetVector<std::string>(block, key)) out.push_back(std::strtod(s.c_str(), nullptr));
  return out;
}

void ParameterInput::CheckRequired(const std::string &block, const std::string &key) const {
  PARTHENON_REQUIRE(DoesParameterExist(block, key),
                    "Parameter file missing required field <" + block + ">/" + key);
}
void ParameterInput::CheckDesired(const std::string &block, const std::string &key) const {
  if (!DoesParameterExist(block, key))
    std::fprintf(stderr, "### WARNING: parameter file missing suggested field <%s>/%s\n",
                 block.c_str(), key.c_str());
}

void ParameterInput::ParameterDump(std::FILE *f) const {
  for (auto &b : order_) {
    std::fprintf(f, "<%s>\n", b.c_str());
    for (auto &kv : blocks_.at(b)) std::fprintf(f, "%s = %s\n", kv.first.c_str(), kv.second.c_str());
    std::fprintf(f, "\n");
  }
}

} // namespace parthenon
