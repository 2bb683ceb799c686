// Minimal JSON reader for the two flat files of the FWI boundary (para_file.json,
// survey_file.json).  Replaces the reference's vendored rapidjson
// (deps/CustomOps/FWI/Src/rapidjson/, used at Parameter.cpp:33-36 and
// Src_Rec.cu:36-38).  Unlike the reference (one getline) it accepts multi-line files.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace fwi {

struct JsonValue {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  double num = 0.0;
  bool is_int = false;  // literal had no '.', 'e' or 'E'
  std::string str;
  std::vector<JsonValue> arr;
  std::vector<std::pair<std::string, JsonValue>> obj;

  const JsonValue *find(const std::string &key) const {
    if (type != Object) return nullptr;
    for (const auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool has(const std::string &key) const { return find(key) != nullptr; }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string &text) : s_(text), p_(0) {}
  JsonValue parse() {
    JsonValue v = value();
    ws();
    if (p_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string &s_;
  size_t p_;
  [[noreturn]] void fail(const char *what) const {
    throw std::runtime_error(std::string("JSON: ") + what + " at offset " + std::to_string(p_));
  }
  void ws() {
    while (p_ < s_.size() && std::isspace(static_cast<unsigned char>(s_[p_]))) ++p_;
  }
  char peek() {
    ws();
    if (p_ >= s_.size()) fail("unexpected end");
    return s_[p_];
  }
  void expect(char c) {
    if (peek() != c) fail("unexpected character");
    ++p_;
  }
  JsonValue value() {
    char c = peek();
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') {
      JsonValue v;
      v.type = JsonValue::String;
      v.str = string();
      return v;
    }
    if (s_.compare(p_, 4, "true") == 0) { p_ += 4; JsonValue v; v.type = JsonValue::Bool; v.b = true; return v; }
    if (s_.compare(p_, 5, "false") == 0) { p_ += 5; JsonValue v; v.type = JsonValue::Bool; v.b = false; return v; }
    if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return JsonValue(); }
    return number();
  }
  JsonValue number() {
    const char *start = s_.c_str() + p_;
    char *end = nullptr;
    double d = std::strtod(start, &end);
    if (end == start) fail("bad number");
    JsonValue v;
    v.type = JsonValue::Number;
    v.num = d;
    v.is_int = true;
    for (const char *q = start; q < end; ++q)
      if (*q == '.' || *q == 'e' || *q == 'E') v.is_int = false;
    p_ += static_cast<size_t>(end - start);
    return v;
  }
  std::string string() {
    expect('"');
    std::string out;
    while (true) {
      if (p_ >= s_.size()) fail("unterminated string");
      char c = s_[p_++];
      if (c == '"') break;
      if (c == '\\') {
        if (p_ >= s_.size()) fail("bad escape");
        char e = s_[p_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {  // keep ASCII range only (paths)
            if (p_ + 4 > s_.size()) fail("bad \\u escape");
            unsigned code = static_cast<unsigned>(std::strtoul(s_.substr(p_, 4).c_str(), nullptr, 16));
            p_ += 4;
            out += static_cast<char>(code & 0x7f);
            break;
          }
          default: out += e;  // \" \\ \/
        }
      } else {
        out += c;
      }
    }
    return out;
  }
  JsonValue array() {
    expect('[');
    JsonValue v;
    v.type = JsonValue::Array;
    if (peek() == ']') { ++p_; return v; }
    while (true) {
      v.arr.push_back(value());
      char c = peek();
      ++p_;
      if (c == ']') break;
      if (c != ',') fail("expected , or ]");
    }
    return v;
  }
  JsonValue object() {
    expect('{');
    JsonValue v;
    v.type = JsonValue::Object;
    if (peek() == '}') { ++p_; return v; }
    while (true) {
      if (peek() != '"') fail("expected key");
      std::string k = string();
      expect(':');
      v.obj.emplace_back(std::move(k), value());
      char c = peek();
      ++p_;
      if (c == '}') break;
      if (c != ',') fail("expected , or }");
    }
    return v;
  }
};

}  // namespace fwi
