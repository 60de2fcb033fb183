#include "config.h"

#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <sstream>

namespace hml {

static std::string strip(const std::string &s) {
  const char *ws = " \t\r\n";
  size_t b = s.find_first_not_of(ws);
  if (b == std::string::npos) return "";
  size_t e = s.find_last_not_of(ws);
  return s.substr(b, e - b + 1);
}

bool CfgFile::load(const std::string &path, std::string &err) {
  std::ifstream f(path);
  if (!f.is_open()) {
    err = "Error opening config file: " + path;
    return false;
  }
  std::string line;
  int lineno = 0;
  while (std::getline(f, line)) {
    ++lineno;
    if (line.empty() || line[0] == '#') continue;  // the reference only treats column-0 '#' as a comment
    size_t eq = line.find('=');
    if (eq == std::string::npos) continue;  // lines without '=' are ignored, as in the reference
    std::string key = strip(line.substr(0, eq)), val = strip(line.substr(eq + 1));
    if (key.empty()) continue;
    char *end = nullptr;
    long v = std::strtol(val.c_str(), &end, 10);  // std::stoi semantics: leading integer, trailing junk ignored
    if (end == val.c_str()) {
      std::ostringstream os;
      os << path << ":" << lineno << ": value of '" << key << "' is not an integer";
      err = os.str();
      return false;
    }
    kv_[key] = static_cast<uint32_t>(v);
  }
  return true;
}

bool CfgFile::get(const std::string &key, uint32_t &out) const {
  auto it = kv_.find(key);
  if (it == kv_.end()) return false;
  out = it->second;
  return true;
}

uint32_t CfgFile::get_or(const std::string &key, uint32_t dflt) const {
  auto it = kv_.find(key);
  return it == kv_.end() ? dflt : it->second;
}

std::string CfgFile::dump() const {
  std::ostringstream os;
  os << "Configuration details are as follow:\n\n*****************************************\n\n";
  for (const auto &p : kv_) {
    os << std::left << std::setw(20) << p.first << " " << std::right << std::setw(20) << p.second << "\n";
  }
  os << "\n*****************************************\n\n";
  return os.str();
}

}  // namespace hml
