// yaggo_lite.hpp -- a tiny getopt_long based option parser used by the hand-written
// stand-ins for the reference's yaggo-generated *_cmdline.hpp headers (yaggo is not
// installed in this image).  TEST INFRASTRUCTURE ONLY (oracle build); written from
// the option specifications in the reference's *.yaggo files, not from generated code.
#ifndef YAGGO_LITE_HPP
#define YAGGO_LITE_HPP
#include <getopt.h>
#include <stdint.h>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace yaggo_lite {

// Streamable object: prints "Error: <msg>" and exits(1) when destroyed.
class error_stream {
  std::ostringstream os_;
  bool active_;
public:
  error_stream() : active_(true) { }
  explicit error_stream(const char* msg) : active_(true) { os_ << msg; }
  error_stream(const error_stream& rhs) : active_(true) { os_ << rhs.os_.str(); const_cast<error_stream&>(rhs).active_ = false; }
  ~error_stream() {
    if(active_) { std::cerr << "Error: " << os_.str() << std::endl; exit(EXIT_FAILURE); }
  }
  template<typename T> error_stream& operator<<(const T& x) { os_ << x; return *this; }
};

enum kind_t { FLAG, U32, U64, U64S, I32, LONG, DOUBLE, CSTR, STRING, CSTR_M, U32_M, U64_M };

struct option_def {
  const char* lng;   // long name or NULL
  char        sht;   // short name or 0
  kind_t      kind;
  void*       dst;
  bool*       given; // may be NULL
  bool        required;
};

inline uint64_t to_u64(const char* s, bool suffix, const char* name) {
  errno = 0;
  char* end = 0;
  while(*s == ' ') ++s;
  if(*s == '-') { error_stream() << "Invalid negative value '" << s << "' for switch " << name; }
  unsigned long long v = strtoull(s, &end, 0);
  if(errno || end == s) { error_stream() << "Invalid numeric value '" << s << "' for switch " << name; }
  if(*end) {
    uint64_t mult = 0;
    if(suffix && end[1] == '\0') {
      switch(*end) {   // SI suffixes, powers of 1000 (doc/jellyfish.man:137-139)
      case 'k': mult = 1000ULL; break;
      case 'M': mult = 1000000ULL; break;
      case 'G': mult = 1000000000ULL; break;
      case 'T': mult = 1000000000000ULL; break;
      case 'P': mult = 1000000000000000ULL; break;
      case 'E': mult = 1000000000000000000ULL; break;
      }
    }
    if(!mult) { error_stream() << "Invalid numeric value '" << s << "' for switch " << name; }
    v *= mult;
  }
  return v;
}

inline long to_long(const char* s, const char* name) {
  errno = 0; char* end = 0;
  long v = strtol(s, &end, 0);
  if(errno || end == s || *end) { error_stream() << "Invalid integer '" << s << "' for switch " << name; }
  return v;
}

inline double to_double(const char* s, const char* name) {
  errno = 0; char* end = 0;
  double v = strtod(s, &end);
  if(errno || end == s || *end) { error_stream() << "Invalid float '" << s << "' for switch " << name; }
  return v;
}

class parser {
  std::vector<option_def> defs_;
  const char*             usage_;
public:
  explicit parser(const char* usage) : usage_(usage) { }
  void add(const char* lng, char sht, kind_t kind, void* dst, bool* given = 0, bool required = false) {
    option_def d = { lng, sht, kind, dst, given, required };
    defs_.push_back(d);
  }

  // Parse; positional arguments are appended to pos.
  void parse(int argc, char* argv[], std::vector<const char*>& pos) {
    std::string           shorts = "+";  // placeholder replaced below
    shorts.clear();
    std::vector<struct option> longs;
    std::vector<bool>     seen(defs_.size(), false);
    for(size_t i = 0; i < defs_.size(); ++i) {
      const option_def& d = defs_[i];
      if(d.sht) { shorts += d.sht; if(d.kind != FLAG) shorts += ':'; }
      if(d.lng) {
        struct option o = { d.lng, d.kind == FLAG ? no_argument : required_argument, 0, (int)(1000 + i) };
        longs.push_back(o);
      }
    }
    struct option h1 = { "help", no_argument, 0, 'h' + 5000 };
    longs.push_back(h1);
    struct option last = { 0, 0, 0, 0 };
    longs.push_back(last);

    optind = 1;
    int c;
    while((c = getopt_long(argc, argv, shorts.c_str(), longs.data(), 0)) != -1) {
      if(c == 'h' + 5000) { std::cout << usage_ << std::endl; exit(0); }
      if(c == '?' || c == ':') { error_stream() << "Invalid command line. " << usage_; }
      size_t idx = defs_.size();
      if(c >= 1000) idx = c - 1000;
      else for(size_t i = 0; i < defs_.size(); ++i) if(defs_[i].sht == c) { idx = i; break; }
      if(idx >= defs_.size()) { error_stream() << "Invalid switch. " << usage_; }
      const option_def& d = defs_[idx];
      std::string name = d.lng ? std::string("--") + d.lng : std::string("-") + d.sht;
      seen[idx] = true;
      if(d.given) *d.given = true;
      switch(d.kind) {
      case FLAG:   *(bool*)d.dst = true; break;
      case U32:    *(uint32_t*)d.dst = (uint32_t)to_u64(optarg, false, name.c_str()); break;
      case U64:    *(uint64_t*)d.dst = to_u64(optarg, false, name.c_str()); break;
      case U64S:   *(uint64_t*)d.dst = to_u64(optarg, true, name.c_str()); break;
      case I32:    *(int32_t*)d.dst = (int32_t)to_long(optarg, name.c_str()); break;
      case LONG:   *(long*)d.dst = (long)to_u64(optarg, false, name.c_str()); break;
      case DOUBLE: *(double*)d.dst = to_double(optarg, name.c_str()); break;
      case CSTR:   *(const char**)d.dst = optarg; break;
      case STRING: *(std::string*)d.dst = optarg; break;
      case CSTR_M: ((std::vector<const char*>*)d.dst)->push_back(optarg); break;
      case U32_M:  ((std::vector<uint32_t>*)d.dst)->push_back((uint32_t)to_u64(optarg, false, name.c_str())); break;
      case U64_M:  ((std::vector<uint64_t>*)d.dst)->push_back(to_u64(optarg, false, name.c_str())); break;
      }
    }
    for(size_t i = 0; i < defs_.size(); ++i)
      if(defs_[i].required && !seen[i]) {
        error_stream() << "Missing required switch "
                       << (defs_[i].lng ? std::string("--") + defs_[i].lng : std::string("-") + defs_[i].sht);
      }
    for(int i = optind; i < argc; ++i) pos.push_back(argv[i]);
  }
};
} // namespace yaggo_lite
#endif
