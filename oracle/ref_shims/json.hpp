// ORACLE -- TEST INFRASTRUCTURE ONLY.  Stand-in for nlohmann/json (un-vendored, absent here) with just the surface the
// compiled reference code touches: object / array access, iteration, implicit conversion of a number to long or double,
// assignment of a double, push_back.  Documents are built in memory by the oracle/ref_wrap_*.cc wrappers; nothing is
// parsed or printed.  Not part of the product.
#pragma once
#include <cstddef>
#include <map>
#include <string>
#include <vector>

namespace nlohmann {
class json {
 public:
  std::map<std::string, json> obj;
  std::vector<json> arr;
  double num = 0;
  long inum = 0;
  bool is_int = false;
  json() {}
  static json integer(long v) { json j; j.inum = v; j.is_int = true; return j; }
  static json real(double v) { json j; j.num = v; return j; }
  json& operator[](const std::string& k) { return obj[k]; }
  const json& operator[](const std::string& k) const { return obj.at(k); }
  json& operator[](const char* k) { return obj[k]; }                       // (a literal key must not reach the built-in long[ptr])
  const json& operator[](const char* k) const { return obj.at(k); }
  json& operator=(double v) { num = v; is_int = false; return *this; }
  operator long() const { return is_int ? inum : (long)num; }
  operator double() const { return is_int ? (double)inum : num; }
  size_t size() const { return arr.size(); }
  const json& at(size_t i) const { return arr.at(i); }
  void push_back(const json& j) { arr.push_back(j); }
  std::vector<json>::const_iterator begin() const { return arr.begin(); }
  std::vector<json>::const_iterator end() const { return arr.end(); }
};
}  // namespace nlohmann
