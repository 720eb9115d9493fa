// parameter_input.hpp — Athena++-style input deck ("<block>" headers, "key = value # note").
// Same accessor names and semantics as the reference's ParameterInput
// (src/parameter_input.hpp; parser src/parameter_input.cpp:100-300): Get* throws if the
// key is absent, GetOrAdd* records the default, command-line overrides are
// "block/key=value" (ModifyFromCmdline, parameter_input.cpp:588-650).
#pragma once
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "types.hpp"

namespace parthenon {

class ParameterInput {
 public:
  ParameterInput() = default;
  void LoadFromString(const std::string &text);
  void LoadFromFile(const std::string &path);
  void ModifyFromCmdline(int argc, char *argv[]);
  void ModifyFromString(const std::string &assignment); // "block/key=value"

  bool DoesBlockExist(const std::string &block) const { return blocks_.count(block) > 0; }
  bool DoesParameterExist(const std::string &block, const std::string &key) const;

  int GetInteger(const std::string &block, const std::string &key) const;
  Real GetReal(const std::string &block, const std::string &key) const;
  bool GetBoolean(const std::string &block, const std::string &key) const;
  std::string GetString(const std::string &block, const std::string &key) const;

  int GetOrAddInteger(const std::string &block, const std::string &key, int def);
  Real GetOrAddReal(const std::string &block, const std::string &key, Real def);
  bool GetOrAddBoolean(const std::string &block, const std::string &key, bool def);
  std::string GetOrAddString(const std::string &block, const std::string &key,
                             const std::string &def);
  std::string GetOrAddString(const std::string &block, const std::string &key,
                             const std::string &def, const std::vector<std::string> &allowed);

  void SetInteger(const std::string &block, const std::string &key, int v);
  void SetReal(const std::string &block, const std::string &key, Real v);
  void SetBoolean(const std::string &block, const std::string &key, bool v);
  void SetString(const std::string &block, const std::string &key, const std::string &v);

  template <typename T>
  std::vector<T> GetVector(const std::string &block, const std::string &key) const;

  void CheckRequired(const std::string &block, const std::string &key) const;
  void CheckDesired(const std::string &block, const std::string &key) const;
  // names of all blocks, in deck order (e.g. to enumerate parthenon/outputN)
  const std::vector<std::string> &BlockNames() const { return order_; }
  void ParameterDump(std::FILE *f) const;

 private:
  const std::string *Find(const std::string &block, const std::string &key) const;
  const std::string &Require(const std::string &block, const std::string &key) const;
  std::map<std::string, std::map<std::string, std::string>> blocks_;
  std::vector<std::string> order_;
};

} // namespace parthenon
