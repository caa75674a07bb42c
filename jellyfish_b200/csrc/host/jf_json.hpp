// jf_json.hpp -- the small JSON subset needed by the jellyfish file header.
// Writer output is byte-compatible with what the reference emits through jsoncpp's
// FastWriter (reference include/jellyfish/generic_file_header.hpp:88-111): no
// whitespace, object keys in sorted order, strings escaped as \" \\ \b \f \n \r \t
// and \u00XX for other control characters, '/' left alone.
#ifndef JFB_JSON_HPP
#define JFB_JSON_HPP
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include <stdexcept>

namespace jfb {

class json {
public:
  enum type_t { NUL, BOOL, INT, UINT, REAL, STR, ARR, OBJ };
private:
  type_t t_;
  bool b_; int64_t i_; uint64_t u_; double d_;
  std::string s_;
  std::vector<json> a_;
  std::map<std::string, json> o_;
public:
  json() : t_(NUL), b_(false), i_(0), u_(0), d_(0) { }
  json(bool b) : t_(BOOL), b_(b), i_(0), u_(0), d_(0) { }
  json(int v) : t_(INT), b_(false), i_(v), u_(0), d_(0) { }
  json(long v) : t_(INT), b_(false), i_(v), u_(0), d_(0) { }
  json(long long v) : t_(INT), b_(false), i_(v), u_(0), d_(0) { }
  json(unsigned v) : t_(UINT), b_(false), i_(0), u_(v), d_(0) { }
  json(unsigned long v) : t_(UINT), b_(false), i_(0), u_(v), d_(0) { }
  json(unsigned long long v) : t_(UINT), b_(false), i_(0), u_(v), d_(0) { }
  json(double v) : t_(REAL), b_(false), i_(0), u_(0), d_(v) { }
  json(const char* s) : t_(STR), b_(false), i_(0), u_(0), d_(0), s_(s) { }
  json(const std::string& s) : t_(STR), b_(false), i_(0), u_(0), d_(0), s_(s) { }
  static json array() { json j; j.t_ = ARR; return j; }
  static json object() { json j; j.t_ = OBJ; return j; }

  type_t type() const { return t_; }
  bool is_null() const { return t_ == NUL; }
  bool has(const std::string& k) const { return t_ == OBJ && o_.count(k); }
  json& operator[](const std::string& k) { if(t_ == NUL) t_ = OBJ; return o_[k]; }
  const json& at(const std::string& k) const {
    static const json null_value;
    if(t_ != OBJ) return null_value;
    std::map<std::string, json>::const_iterator it = o_.find(k);
    return it == o_.end() ? null_value : it->second;
  }
  void erase(const std::string& k) { o_.erase(k); }
  void push_back(const json& v) { if(t_ == NUL) t_ = ARR; a_.push_back(v); }
  size_t size() const { return t_ == ARR ? a_.size() : (t_ == OBJ ? o_.size() : 0); }
  const json& operator[](size_t i) const { return a_[i]; }
  const std::map<std::string, json>& members() const { return o_; }

  uint64_t as_u64(uint64_t dflt = 0) const {
    switch(t_) { case INT: return (uint64_t)i_; case UINT: return u_; case REAL: return (uint64_t)d_; case BOOL: return b_; default: return dflt; }
  }
  int64_t as_i64(int64_t dflt = 0) const {
    switch(t_) { case INT: return i_; case UINT: return (int64_t)u_; case REAL: return (int64_t)d_; case BOOL: return b_; default: return dflt; }
  }
  double as_double(double dflt = 0) const {
    switch(t_) { case INT: return (double)i_; case UINT: return (double)u_; case REAL: return d_; default: return dflt; }
  }
  bool as_bool(bool dflt = false) const {
    switch(t_) { case BOOL: return b_; case INT: return i_ != 0; case UINT: return u_ != 0; default: return dflt; }
  }
  std::string as_string(const std::string& dflt = "") const { return t_ == STR ? s_ : dflt; }

  // ---- writer -------------------------------------------------------------
  static void quote(std::string& out, const std::string& s) {
    out += '"';
    for(size_t i = 0; i < s.size(); ++i) {
      unsigned char c = (unsigned char)s[i];
      switch(c) {
      case '"':  out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\b': out += "\\b"; break;
      case '\f': out += "\\f"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      default:
        if(c > 0 && c <= 0x1f) { char buf[8]; snprintf(buf, sizeof(buf), "\\u%04X", c); out += buf; }
        else out += (char)c;
      }
    }
    out += '"';
  }
  static std::string real_to_string(double v) {
    // jsoncpp 0.6: "%#.16g", then strip trailing zeros but keep one digit after the point
    char buf[64];
    snprintf(buf, sizeof(buf), "%#.16g", v);
    char* ch = buf + strlen(buf) - 1;
    if(*ch != '0') return buf;
    while(ch > buf && *ch == '0') --ch;
    char* last_nonzero = ch;
    while(ch >= buf) {
      switch(*ch) {
      case '0': case '1': case '2': case '3': case '4': case '5': case '6': case '7': case '8': case '9':
        --ch; continue;
      case '.':
        *(last_nonzero + 2) = '\0';   // keep "x.0"
        return buf;
      default:
        return buf;
      }
    }
    return buf;
  }
  void write(std::string& out) const {
    char buf[32];
    switch(t_) {
    case NUL:  out += "null"; break;
    case BOOL: out += b_ ? "true" : "false"; break;
    case INT:  snprintf(buf, sizeof(buf), "%lld", (long long)i_); out += buf; break;
    case UINT: snprintf(buf, sizeof(buf), "%llu", (unsigned long long)u_); out += buf; break;
    case REAL: out += real_to_string(d_); break;
    case STR:  quote(out, s_); break;
    case ARR:
      out += '[';
      for(size_t i = 0; i < a_.size(); ++i) { if(i) out += ','; a_[i].write(out); }
      out += ']';
      break;
    case OBJ: {
      out += '{';
      bool first = true;
      for(std::map<std::string, json>::const_iterator it = o_.begin(); it != o_.end(); ++it) {
        if(!first) out += ',';
        first = false;
        quote(out, it->first); out += ':'; it->second.write(out);
      }
      out += '}';
      break; }
    }
  }
  std::string dump() const { std::string s; write(s); return s; }

  // ---- reader -------------------------------------------------------------
  static bool parse(const char* begin, const char* end, json& out) {
    const char* p = begin;
    if(!parse_value(p, end, out)) return false;
    skip_ws(p, end);
    return p == end;
  }
private:
  static void skip_ws(const char*& p, const char* e) { while(p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
  static bool parse_string(const char*& p, const char* e, std::string& s) {
    if(p >= e || *p != '"') return false;
    ++p; s.clear();
    while(p < e && *p != '"') {
      if(*p == '\\') {
        if(++p >= e) return false;
        switch(*p) {
        case '"': s += '"'; break;  case '\\': s += '\\'; break; case '/': s += '/'; break;
        case 'b': s += '\b'; break; case 'f': s += '\f'; break; case 'n': s += '\n'; break;
        case 'r': s += '\r'; break; case 't': s += '\t'; break;
        case 'u': {
          if(e - p < 5) return false;
          char hex[5] = { p[1], p[2], p[3], p[4], 0 };
          unsigned cp = (unsigned)strtoul(hex, 0, 16);
          p += 4;
          if(cp < 0x80) s += (char)cp;
          else if(cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
          else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
          break; }
        default: return false;
        }
        ++p;
      } else s += *p++;
    }
    if(p >= e) return false;
    ++p;
    return true;
  }
  static bool parse_value(const char*& p, const char* e, json& out) {
    skip_ws(p, e);
    if(p >= e) return false;
    if(*p == '{') {
      ++p; out = object();
      skip_ws(p, e);
      if(p < e && *p == '}') { ++p; return true; }
      for(;;) {
        skip_ws(p, e);
        std::string key;
        if(!parse_string(p, e, key)) return false;
        skip_ws(p, e);
        if(p >= e || *p != ':') return false;
        ++p;
        json v;
        if(!parse_value(p, e, v)) return false;
        out.o_[key] = v;
        skip_ws(p, e);
        if(p < e && *p == ',') { ++p; continue; }
        if(p < e && *p == '}') { ++p; return true; }
        return false;
      }
    }
    if(*p == '[') {
      ++p; out = array();
      skip_ws(p, e);
      if(p < e && *p == ']') { ++p; return true; }
      for(;;) {
        json v;
        if(!parse_value(p, e, v)) return false;
        out.a_.push_back(v);
        skip_ws(p, e);
        if(p < e && *p == ',') { ++p; continue; }
        if(p < e && *p == ']') { ++p; return true; }
        return false;
      }
    }
    if(*p == '"') { std::string s; if(!parse_string(p, e, s)) return false; out = json(s); return true; }
    if(e - p >= 4 && !strncmp(p, "true", 4)) { p += 4; out = json(true); return true; }
    if(e - p >= 5 && !strncmp(p, "false", 5)) { p += 5; out = json(false); return true; }
    if(e - p >= 4 && !strncmp(p, "null", 4)) { p += 4; out = json(); return true; }
    // number
    const char* s = p;
    bool neg = false, real = false;
    if(p < e && *p == '-') { neg = true; ++p; }
    while(p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) {
      if(*p == '.' || *p == 'e' || *p == 'E') real = true;
      ++p;
    }
    if(p == s || (neg && p == s + 1)) return false;
    std::string num(s, p);
    if(real) out = json(strtod(num.c_str(), 0));
    else if(neg) out = json((long long)strtoll(num.c_str(), 0, 10));
    else out = json((unsigned long long)strtoull(num.c_str(), 0, 10));
    return true;
  }
};

} // namespace jfb
#endif
