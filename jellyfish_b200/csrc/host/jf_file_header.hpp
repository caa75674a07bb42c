// jf_file_header.hpp -- the jellyfish database header.
// Format contract (reference include/jellyfish/generic_file_header.hpp:88-152 and
// include/jellyfish/file_header.hpp:26-108):
//   9 decimal digits = length of (terse JSON + padding), the JSON, then '\0' padding so
//   that 9 + length is a multiple of "alignment" (8).  The record body starts there.
#ifndef JFB_FILE_HEADER_HPP
#define JFB_FILE_HEADER_HPP
#include <unistd.h>
#include <limits.h>
#include <sys/utsname.h>
#include <ctime>
#include <cctype>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "jf_json.hpp"
#include "jf_matrix.hpp"

namespace jfb {

class file_header {
  static const int MAX_HEADER_DIGITS = 9;
  json   root_;
  size_t offset_;

  static void chomp(std::string& s) {
    size_t found = s.find_last_not_of(" \t\f\v\n\r");
    if(found != std::string::npos) s.erase(found + 1); else s.clear();
  }
public:
  file_header() : root_(json::object()), offset_(0) { root_["alignment"] = json(8); }

  const json& root() const { return root_; }
  json& root() { return root_; }
  size_t offset() const { return offset_; }
  int alignment() const { int a = (int)root_.at("alignment").as_i64(0); return a > 0 ? a : 0; }

  // -- generic part ---------------------------------------------------------
  void fill_standard() {
    const char* sde = getenv("SOURCE_DATE_EPOCH");
    // hostname
    if(sde) root_["hostname"] = json("hostname");
    else { struct utsname u; root_["hostname"] = json(uname(&u) == -1 ? "" : u.nodename); }
    // pwd
    if(sde) root_["pwd"] = json(".");
    else { char path[PATH_MAX + 1]; if(!getcwd(path, sizeof(path))) path[0] = '\0'; root_["pwd"] = json(path); }
    // time
    time_t t = time(0);
    std::string ts;
    if(sde) {
      std::istringstream iss(sde);
      iss >> t;
      if(iss.fail() || !iss.eof()) { std::cerr << "Error: Cannot parse SOURCE_DATE_EPOCH as integer\n"; exit(27); }
      ts = asctime(gmtime(&t));
    } else ts = ctime(&t);
    chomp(ts);
    root_["time"] = json(ts);
    // exe_path
    char path[PATH_MAX + 1];
    ssize_t l = readlink("/proc/self/exe", path, sizeof(path));
    root_["exe_path"] = json(l == -1 ? std::string() : std::string(path, l));
  }
  void set_cmdline(int argc, char* argv[]) {
    json a = json::array();
    for(int i = 0; i < argc; ++i) a.push_back(json(argv[i]));
    root_["cmdline"] = a;
  }
  std::vector<std::string> cmdline() const {
    std::vector<std::string> res;
    const json& a = root_.at("cmdline");
    for(size_t i = 0; i < a.size(); ++i) res.push_back(a[i].as_string());
    return res;
  }

  void write(std::ostream& os) {
    std::string h = root_.dump();
    size_t hlen = h.size();
    int align = alignment(), padding = 0;
    if(align > 0) {
      padding = (MAX_HEADER_DIGITS + h.size()) % align;
      if(padding) hlen += align - padding;
    }
    char len[16];
    snprintf(len, sizeof(len), "%09lu", (unsigned long)hlen);
    os.write(len, MAX_HEADER_DIGITS);
    os.write(h.data(), h.size());
    offset_ = MAX_HEADER_DIGITS + hlen;
    if(padding) { std::string pad(align - padding, '\0'); os.write(pad.data(), pad.size()); }
  }

  bool read(std::istream& is) {
    std::string len;
    for(int i = 0; i < MAX_HEADER_DIGITS && isdigit(is.peek()); ++i) len += (char)is.get();
    if(is.peek() != '{') return false;
    unsigned long hlen = strtoul(len.c_str(), 0, 10);
    if(hlen < 2) return false;
    offset_ = MAX_HEADER_DIGITS + hlen;
    std::vector<char> buf(hlen);
    is.read(buf.data(), hlen);
    if(!is.good()) return false;
    const char* end = buf.data() + hlen;
    while(end > buf.data() && *(end - 1) == '\0') --end;
    return json::parse(buf.data(), end, root_);
  }
  // parse from memory (mmap'd database); returns false on failure
  bool read(const char* data, size_t size) {
    size_t i = 0; std::string len;
    for(; i < (size_t)MAX_HEADER_DIGITS && i < size && isdigit((unsigned char)data[i]); ++i) len += data[i];
    if(i >= size || data[i] != '{') return false;
    unsigned long hlen = strtoul(len.c_str(), 0, 10);
    if(hlen < 2 || i + hlen > size) return false;
    offset_ = MAX_HEADER_DIGITS + hlen;
    const char* end = data + i + hlen;
    while(end > data + i && *(end - 1) == '\0') --end;
    return json::parse(data + i, end, root_);
  }

  // -- table description ------------------------------------------------------
  uint64_t size() const { return root_.at("size").as_u64(); }
  void size(uint64_t s) { root_["size"] = json((unsigned long long)s); }
  unsigned key_len() const { return (unsigned)root_.at("key_len").as_u64(); }
  void key_len(unsigned k) { root_["key_len"] = json(k); }
  unsigned val_len() const { return (unsigned)root_.at("val_len").as_u64(); }
  void val_len(unsigned k) { root_["val_len"] = json(k); }
  unsigned max_reprobe() const { return (unsigned)root_.at("max_reprobe").as_u64(); }
  void max_reprobe(unsigned m) { root_["max_reprobe"] = json(m); }
  unsigned counter_len() const { return (unsigned)root_.at("counter_len").as_u64(); }
  void counter_len(unsigned l) { root_["counter_len"] = json(l); }
  std::string format() const { return root_.at("format").as_string(); }
  void format(const std::string& s) { root_["format"] = json(s); }
  bool canonical() const { return root_.at("canonical").as_bool(false); }
  void canonical(bool v) { root_["canonical"] = json(v); }
  double fpr() const { return root_.at("fpr").as_double(); }
  void fpr(double f) { root_["fpr"] = json(f); }
  unsigned long nb_hashes() const { return (unsigned long)root_.at("nb_hashes").as_u64(); }
  void nb_hashes(unsigned long n) { root_["nb_hashes"] = json((unsigned)n); }

  // max_reprobe() must be set first; reprobes has max_reprobe()+1 entries
  void set_reprobes(const uint64_t* reprobes) {
    json a = json::array();
    for(unsigned i = 0; i <= max_reprobe(); ++i) a.push_back(json((unsigned long long)reprobes[i]));
    root_["reprobes"] = a;
  }
  std::vector<uint64_t> reprobes() const {
    std::vector<uint64_t> r;
    const json& a = root_.at("reprobes");
    for(size_t i = 0; i < a.size(); ++i) r.push_back(a[i].as_u64());
    return r;
  }
  uint64_t max_reprobe_offset() const {
    const json& a = root_.at("reprobes");
    return max_reprobe() < a.size() ? a[max_reprobe()].as_u64() : 0;
  }

  void matrix(const gf2_matrix& m, int i = 1) {
    std::string name = "matrix" + std::to_string(i);
    json o = json::object();
    o["r"] = json(m.r());
    o["c"] = json(m.c());
    if(m.is_low_identity()) {
      o["identity"] = json(true);
    } else {
      o["identity"] = json(false);
      json cols = json::array();
      for(unsigned j = 0; j < m.c(); ++j) cols.push_back(json((unsigned long long)m[j]));
      o["columns"] = cols;
    }
    root_[name] = o;
  }
  gf2_matrix matrix(int i = 1) const {
    std::string name = "matrix" + std::to_string(i);
    const json& o = root_.at(name);
    unsigned r = (unsigned)o.at("r").as_u64(), c = (unsigned)o.at("c").as_u64();
    if(o.at("identity").as_bool(false))
      return r == c ? gf2_matrix::identity(c) : gf2_matrix::low_identity(r, c);
    std::vector<uint64_t> raw(c, 0);
    const json& cols = o.at("columns");
    for(unsigned j = 0; j < c && j < cols.size(); ++j) raw[j] = cols[j].as_u64();
    return gf2_matrix(r, c, raw.begin());
  }
};

} // namespace jfb
#endif
