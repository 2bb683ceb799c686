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
  unsigned hex4() {
    if (p_ + 4 > s_.size()) fail("bad \\u escape");
    unsigned code = 0;
    for (int k = 0; k < 4; k++) {
      const char h = s_[p_++];
      code <<= 4;
      if (h >= '0' && h <= '9') code |= static_cast<unsigned>(h - '0');
      else if (h >= 'a' && h <= 'f') code |= static_cast<unsigned>(h - 'a' + 10);
      else if (h >= 'A' && h <= 'F') code |= static_cast<unsigned>(h - 'A' + 10);
      else fail("bad \\u escape");
    }
    return code;
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
          case 'u': {  // \uXXXX (with surrogate pairs) -> UTF-8, like rapidjson in the reference
            unsigned code = hex4();
            if (code >= 0xD800 && code <= 0xDBFF) {   // high surrogate: a low one must follow
              if (p_ + 2 <= s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
                p_ += 2;
                const unsigned lo = hex4();
                if (lo < 0xDC00 || lo > 0xDFFF) fail("bad surrogate pair");
                code = 0x10000 + ((code - 0xD800) << 10) + (lo - 0xDC00);
              } else {
                fail("lone surrogate");
              }
            } else if (code >= 0xDC00 && code <= 0xDFFF) {
              fail("lone surrogate");
            }
            if (code < 0x80) {
              out += static_cast<char>(code);
            } else if (code < 0x800) {
              out += static_cast<char>(0xC0 | (code >> 6));
              out += static_cast<char>(0x80 | (code & 0x3F));
            } else if (code < 0x10000) {
              out += static_cast<char>(0xE0 | (code >> 12));
              out += static_cast<char>(0x80 | ((code >> 6) & 0x3F));
              out += static_cast<char>(0x80 | (code & 0x3F));
            } else {
              out += static_cast<char>(0xF0 | (code >> 18));
              out += static_cast<char>(0x80 | ((code >> 12) & 0x3F));
              out += static_cast<char>(0x80 | ((code >> 6) & 0x3F));
              out += static_cast<char>(0x80 | (code & 0x3F));
            }
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
