// Minimal JSON reader/writer for the pilotguru wire formats: the reference uses nlohmann/json
// 2.1.1 (un-vendored; docker/Dockerfile:34) through ReadJsonFile / JsonWriteTimestampedRealData / WriteJsonFile
// (src/io/json_converters.cc:172-202).  Reading is a pull parser that streams "array of flat objects" tables
// straight into columns (an hour of 500 Hz IMU is ~1.8 M records per file); writing reproduces dump(2):
// two-space indent, object keys in alphabetical order (nlohmann's default std::map), integers without a decimal
// point.  Doubles are printed with 17 significant digits so that parsing the file returns the exact values.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <charconv>
#include <string_view>
#include <algorithm>

#include "check.hpp"

namespace pgbhost {

class JsonReader {
 public:
  explicit JsonReader(const std::string& text) : s_(text), p_(s_.data()), end_(s_.data() + s_.size()) {}
  static std::string Slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    PGB_CHECK(f != nullptr) << "cannot open JSON file " << path;
    std::string out;
    if (fseek(f, 0, SEEK_END) == 0) {
      const long n = ftell(f);
      if (n > 0) out.resize((size_t)n);
      rewind(f);
    }
    size_t got = out.empty() ? 0 : fread(&out[0], 1, out.size(), f);
    if (got < out.size()) out.resize(got);
    char buf[1 << 16];                                  // whatever follows (or everything, for an unseekable stream)
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, got);
    fclose(f);
    return out;
  }
  void SkipWs() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_; }
  char Peek() { SkipWs(); PGB_CHECK(p_ < end_) << "unexpected end of JSON"; return *p_; }
  void Expect(char c) { PGB_CHECK(Peek() == c) << "JSON parse error: expected '" << c << "' at offset " << (p_ - s_.data()); ++p_; }
  bool TryConsume(char c) { if (Peek() == c) { ++p_; return true; } return false; }
  std::string String() {
    Expect('"');
    std::string out;
    while (p_ < end_ && *p_ != '"') {
      if (*p_ == '\\' && p_ + 1 < end_) {
        ++p_;
        switch (*p_) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'r': out.push_back('\r'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case 'u': out.push_back('?'); p_ += 4; break;  // field names and values of this wire format are ASCII
          default: out.push_back(*p_);
        }
        ++p_;
      } else {
        out.push_back(*p_++);
      }
    }
    PGB_CHECK(p_ < end_) << "unterminated JSON string";
    ++p_;
    return out;
  }
  // An object key as a view into the text (the common case: no escapes); falls back to String() otherwise.
  std::string_view Key(std::string* scratch) {
    SkipWs();
    PGB_CHECK(p_ < end_ && *p_ == '"') << "JSON parse error: expected '\"' at offset " << (p_ - s_.data());
    const char* b = p_ + 1;
    const char* q = b;
    while (q < end_ && *q != '"' && *q != '\\') ++q;
    if (q < end_ && *q == '"') { p_ = q + 1; return std::string_view(b, (size_t)(q - b)); }
    *scratch = String();
    return *scratch;
  }
  // A number, kept exact when it is an integer literal.
  void Number(double* d, int64_t* i, bool* is_int) {
    SkipWs();
    const char* b = p_;
    bool integral = true;
    if (p_ < end_ && (*p_ == '-' || *p_ == '+')) ++p_;
    while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) {
      if (*p_ == '.' || *p_ == 'e' || *p_ == 'E') integral = false;
      ++p_;
    }
    PGB_CHECK(p_ > b) << "JSON parse error: number expected at offset " << (b - s_.data());
    // std::from_chars: correctly rounded like strtod, no locale, no copy; it does not take a leading '+'
    const char* nb = (*b == '+') ? b + 1 : b;
    *is_int = integral;
    if (integral) {
      const auto r = std::from_chars(nb, p_, *i);
      if (r.ec == std::errc() && r.ptr == p_) { *d = (double)*i; return; }
      *is_int = false;                                   // out of the int64 range: keep it as a double
    }
    const auto r = std::from_chars(nb, p_, *d);
    PGB_CHECK(r.ec == std::errc() || r.ec == std::errc::result_out_of_range)
        << "JSON parse error: malformed number at offset " << (b - s_.data());
    if (r.ec == std::errc::result_out_of_range) *d = strtod(std::string(nb, p_).c_str(), nullptr);  // +-inf / denormal, like strtod
    *i = (int64_t)*d;
  }
  double Double() { double d; int64_t i; bool b; Number(&d, &i, &b); return b ? (double)i : d; }
  int64_t Int() { double d; int64_t i; bool b; Number(&d, &i, &b); return i; }
  bool Bool() {
    SkipWs();
    if (end_ - p_ >= 4 && !strncmp(p_, "true", 4)) { p_ += 4; return true; }
    PGB_CHECK(end_ - p_ >= 5 && !strncmp(p_, "false", 5)) << "JSON parse error: boolean expected";
    p_ += 5;
    return false;
  }
  void SkipValue() {
    const char c = Peek();
    if (c == '"') { String(); return; }
    if (c == '{') {
      ++p_;
      if (TryConsume('}')) return;
      do { String(); Expect(':'); SkipValue(); } while (TryConsume(','));
      Expect('}');
      return;
    }
    if (c == '[') {
      ++p_;
      if (TryConsume(']')) return;
      do { SkipValue(); } while (TryConsume(','));
      Expect(']');
      return;
    }
    if (c == 't' || c == 'f') { Bool(); return; }
    if (c == 'n') { PGB_CHECK(end_ - p_ >= 4) << "JSON parse error"; p_ += 4; return; }
    Double();
  }
  // Positions the reader at the value of `key` of the top-level object; false if absent.
  bool FindTopLevel(const std::string& key) {
    p_ = s_.data();
    Expect('{');
    if (TryConsume('}')) return false;
    do {
      const std::string k = String();
      Expect(':');
      if (k == key) return true;
      SkipValue();
    } while (TryConsume(','));
    return false;
  }

 private:
  const std::string& s_;
  const char* p_;
  const char* end_;
};

// root[table] = [ {field: number, ...}, ... ]  ->  one column per requested field (doubles) and one int64 column.
// Missing fields are fatal, like nlohmann's implicit conversion of a null (type_error -> abort).
struct Table {
  std::vector<std::vector<double>> real;  // [field][row]
  std::vector<int64_t> integer;           // the integer field (time_usec)
  size_t rows() const { return integer.size(); }
};

inline Table ReadTable(const std::string& path, const std::string& table, const std::vector<std::string>& real_fields,
                       const std::string& int_field) {
  const std::string text = JsonReader::Slurp(path);
  JsonReader r(text);
  PGB_CHECK(r.FindTopLevel(table)) << path << ": no \"" << table << "\" element";
  Table t;
  t.real.resize(real_fields.size());
  std::string scratch;
  std::vector<char> seen(real_fields.size() + 1, 0);
  r.Expect('[');
  if (!r.TryConsume(']')) {
    do {
      r.Expect('{');
      std::fill(seen.begin(), seen.end(), 0);
      if (!r.TryConsume('}')) {
        do {
          const std::string_view k = r.Key(&scratch);
          r.Expect(':');
          bool used = false;
          for (size_t f = 0; f < real_fields.size() && !used; f++)
            if (k == real_fields[f]) { t.real[f].push_back(r.Double()); seen[f] = 1; used = true; }
          if (!used && k == int_field) { t.integer.push_back(r.Int()); seen.back() = 1; used = true; }
          if (!used) r.SkipValue();
        } while (r.TryConsume(','));
        r.Expect('}');
      }
      for (size_t f = 0; f < seen.size(); f++)
        PGB_CHECK(seen[f]) << path << ": record " << t.integer.size() << " of \"" << table << "\" lacks field \""
                           << (f < real_fields.size() ? real_fields[f] : int_field) << "\"";
    } while (r.TryConsume(','));
    r.Expect(']');
  }
  PGB_CHECK(t.rows() > 0) << path << ": \"" << table << "\" is empty";  // CHECK(!entries_list.empty()), fit_motion.cc:113,128
  return t;
}

// "%.17g" (std::to_chars with chars_format::general and precision 17 is specified as exactly that), plus ".0" for
// integral values the way nlohmann prints doubles.  Returns the number of characters written to buf (>= 40 bytes).
// Non-finite values become `null`, as nlohmann::json::dump writes them (the reference produces them: the acos of a
// rotation cosine that rounds above 1 in Projected2DDirectionsToTurnAngles, horizontal_flatten.cc:56-61, is NaN).
// Digits: the reference pins nlohmann/json 2.1.1 (docker/Dockerfile:34), which prints 15 significant digits; this writer
// prints 17 so that a value survives the round trip -- compare numerically, not textually.
inline size_t FormatDoubleTo(char* buf, double v) {
  if (!std::isfinite(v)) { memcpy(buf, "null", 5); return 4; }
  char* e = std::to_chars(buf, buf + 32, v, std::chars_format::general, 17).ptr;
  bool plain = true;
  for (const char* c = buf; c < e; ++c)
    if (*c == '.' || *c == 'e' || *c == 'E') { plain = false; break; }
  if (plain) { *e++ = '.'; *e++ = '0'; }
  *e = 0;
  return (size_t)(e - buf);
}
inline std::string FormatDouble(double v) {
  char buf[40];
  const size_t n = FormatDoubleTo(buf, v);
  return std::string(buf, n);
}

// JsonWriteTimestampedRealData (json_converters.cc:184-202): {root: [{"time_usec": t, value_name: v}, ...]} with
// keys in alphabetical order inside each record.
inline void JsonWriteTimestampedRealData(const std::vector<int64_t>& times_usec, const std::vector<double>& values,
                                         const std::string& filename, const std::string& root_element_name,
                                         const std::string& value_name) {
  PGB_CHECK(times_usec.size() == values.size()) << "times/values size mismatch";  // CHECK_EQ
  FILE* f = fopen(filename.c_str(), "w");
  PGB_CHECK(f != nullptr) << "cannot write " << filename;
  const std::string time_name = "time_usec";
  const bool value_first = value_name < time_name;
  if (values.empty()) {
    fprintf(f, "{\n  \"%s\": null\n}\n", root_element_name.c_str());  // out_json[root] = {} dumps as null
  } else {
    fprintf(f, "{\n  \"%s\": [\n", root_element_name.c_str());
    std::string out;
    out.reserve(1 << 20);
    char num[40];
    for (size_t i = 0; i < values.size(); i++) {
      const size_t nv = FormatDoubleTo(num, values[i]);
      char tnum[24];
      const size_t nt = (size_t)(std::to_chars(tnum, tnum + sizeof tnum, (long long)times_usec[i]).ptr - tnum);
      out += "    {\n      \"";
      if (value_first) {
        out += value_name; out += "\": "; out.append(num, nv); out += ",\n      \"time_usec\": "; out.append(tnum, nt);
      } else {
        out += "time_usec\": "; out.append(tnum, nt); out += ",\n      \""; out += value_name; out += "\": "; out.append(num, nv);
      }
      out += i + 1 < values.size() ? "\n    },\n" : "\n    }\n";
      if (out.size() > (1 << 20) - 256) { fwrite(out.data(), 1, out.size(), f); out.clear(); }
    }
    fwrite(out.data(), 1, out.size(), f);
    fprintf(f, "  ]\n}\n");
  }
  fclose(f);
}

}  // namespace pgbhost
