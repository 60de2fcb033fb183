// Parameter-file reader for the reference's `key = value` .cfg format.
// Replaces reference include/Config.h:6-25 + src/Config.cpp:4-52 (same accepted syntax: '#' comment
// lines, blank lines, whitespace around key and value, integer values); unlike the reference a missing
// file or key is reported as an error instead of an uncaught exception.
#pragma once
#include <cstdint>
#include <map>
#include <string>

namespace hml {

class CfgFile {
 public:
  // Returns false (and sets err) when the file cannot be opened or a value is not an integer.
  bool load(const std::string &path, std::string &err);
  bool has(const std::string &key) const { return kv_.count(key) != 0; }
  // Reference semantics: a missing key is fatal (Config.h:14-20 throws); here it yields false.
  bool get(const std::string &key, uint32_t &out) const;
  uint32_t get_or(const std::string &key, uint32_t dflt) const;
  void set(const std::string &key, uint32_t v) { kv_[key] = v; }  // reference Config::setValue
  // Same dump layout as the reference prints at start-up (src/Config.cpp:40-51): keys sorted,
  // key left-aligned in 20 columns, value right-aligned in 20 columns.
  std::string dump() const;

 private:
  std::map<std::string, uint32_t> kv_;
};

}  // namespace hml
