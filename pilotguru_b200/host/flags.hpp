// gflags-compatible command line for the host binaries: --name=value, --name value, --boolflag / --noboolflag,
// single or double dash; unknown flags are fatal (gflags: "ERROR: unknown command line flag").
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>

namespace pgbhost {

class Flags {
 public:
  void String(const char* name, std::string* v, const char* help) { reg_[name] = {kString, v, help}; }
  void Int64(const char* name, int64_t* v, const char* help) { reg_[name] = {kInt, v, help}; }
  void Double(const char* name, double* v, const char* help) { reg_[name] = {kDouble, v, help}; }
  void Bool(const char* name, bool* v, const char* help) { reg_[name] = {kBool, v, help}; }

  void Parse(int argc, char** argv) {
    for (int i = 1; i < argc; i++) {
      const char* a = argv[i];
      if (a[0] != '-') continue;  // positional arguments are ignored, as the reference binaries take none
      a += (a[1] == '-') ? 2 : 1;
      if (!*a) continue;
      std::string name(a), value;
      bool has_value = false;
      const size_t eq = name.find('=');
      if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); has_value = true; }
      if (name == "help") { Usage(argv[0]); exit(1); }
      if (name == "logtostderr" || name == "alsologtostderr" || name == "v" || name == "minloglevel") {  // glog's own
        if (!has_value && name != "logtostderr" && name != "alsologtostderr" && i + 1 < argc) i++;
        verbose = verbose || name == "logtostderr" || name == "alsologtostderr";
        continue;
      }
      auto it = reg_.find(name);
      if (it == reg_.end() && name.rfind("no", 0) == 0) {
        auto nb = reg_.find(name.substr(2));
        if (nb != reg_.end() && nb->second.type == kBool && !has_value) { *(bool*)nb->second.ptr = false; continue; }
      }
      if (it == reg_.end()) {
        std::cerr << "ERROR: unknown command line flag '" << name << "'" << std::endl;
        exit(1);
      }
      Entry& e = it->second;
      if (e.type == kBool) {
        if (!has_value) { *(bool*)e.ptr = true; continue; }
        *(bool*)e.ptr = (value == "true" || value == "1" || value == "t" || value == "yes" || value == "y");
        continue;
      }
      if (!has_value) {
        if (i + 1 >= argc) { std::cerr << "ERROR: flag '" << name << "' is missing its argument" << std::endl; exit(1); }
        value = argv[++i];
      }
      char* endp = nullptr;
      if (e.type == kString) *(std::string*)e.ptr = value;
      if (e.type == kInt) {
        *(int64_t*)e.ptr = strtoll(value.c_str(), &endp, 0);
        if (!*value.c_str() || *endp) { std::cerr << "ERROR: illegal value '" << value << "' specified for int64 flag '" << name << "'" << std::endl; exit(1); }
      }
      if (e.type == kDouble) {
        *(double*)e.ptr = strtod(value.c_str(), &endp);
        if (!*value.c_str() || *endp) { std::cerr << "ERROR: illegal value '" << value << "' specified for double flag '" << name << "'" << std::endl; exit(1); }
      }
    }
  }
  void Usage(const char* prog) {
    std::cerr << prog << ":\n";
    for (auto& kv : reg_) std::cerr << "    -" << kv.first << " (" << kv.second.help << ")\n";
  }
  bool verbose = false;  // --logtostderr: LOG(INFO) lines go to stderr

 private:
  enum Type { kString, kInt, kDouble, kBool };
  struct Entry { Type type; void* ptr; const char* help; };
  std::map<std::string, Entry> reg_;
};

}  // namespace pgbhost
