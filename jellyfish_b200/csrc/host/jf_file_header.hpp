// jf_file_header.hpp -- the jellyfish database header.
// Format contract (reference include/jellyfish/generic_file_header.hpp:88-152 and
// include/jellyfish/file_header.hpp:26-108):
//   9 decimal digits = length of (terse JSON + padding), the JSON, then '\0' padding so
//   that 9 + length is a multiple of "alignment" (8).  The record body starts there.
// The byte layout is the format's; the code is this repository's own (framing as a pure string function, provenance
// fields as separate helpers).
#ifndef JFB_FILE_HEADER_HPP
#define JFB_FILE_HEADER_HPP
#include <unistd.h>
#include <limits.h>
#include <sys/utsname.h>
#include <cstdlib>
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
  // Framing of the header: LEN_DIGITS decimal digits giving the number of bytes that follow up to the record body, the terse
  // JSON text, zero bytes up to the next multiple of "alignment" (counted from the start of the file).
  static constexpr size_t LEN_DIGITS = 9;
  json   root_;
  size_t offset_;

  static std::string frame(const std::string& text, size_t align) {
    size_t total = LEN_DIGITS + text.size();
    if(align > 1 && total % align) total += align - total % align;
    char digits[LEN_DIGITS + 8];
    snprintf(digits, sizeof(digits), "%0*zu", (int)LEN_DIGITS, total - LEN_DIGITS);
    std::string out(digits, LEN_DIGITS);
    out += text;
    out.resize(total, '\0');
    return out;
  }
  // [*lo, *hi) = the JSON text inside a framed header at `p`; false when `p` does not start with one
  static bool unframe(const char* p, size_t n, size_t* lo, size_t* hi, size_t* body) {
    size_t d = 0; unsigned long len = 0;
    while(d < LEN_DIGITS && d < n && p[d] >= '0' && p[d] <= '9') { len = len * 10 + (unsigned long)(p[d] - '0'); ++d; }
    if(d >= n || p[d] != '{' || len < 2 || d + len > n) return false;
    size_t e = d + len;
    while(e > d && p[e - 1] == '\0') --e;
    *lo = d; *hi = e; *body = LEN_DIGITS + len;
    return true;
  }
  // provenance fields ("hostname", "pwd", "time", "exe_path"); SOURCE_DATE_EPOCH pins them for reproducible files
  static std::string host_name(bool pinned) {
    if(pinned) return "hostname";
    struct utsname u;
    return uname(&u) == 0 ? std::string(u.nodename) : std::string();
  }
  static std::string work_dir(bool pinned) {
    if(pinned) return ".";
    std::vector<char> buf(PATH_MAX + 1, '\0');
    return getcwd(buf.data(), buf.size()) ? std::string(buf.data()) : std::string();
  }
  static std::string time_stamp(const char* epoch) {
    time_t t = time(nullptr);
    const char* text;
    if(epoch) {
      char* end = nullptr;
      const long long v = strtoll(epoch, &end, 10);
      if(end == epoch || *end != '\0') { std::cerr << "SOURCE_DATE_EPOCH is not an integer" << std::endl; exit(1); }
      t = (time_t)v;
      text = asctime(gmtime(&t));
    } else text = ctime(&t);
    std::string s(text ? text : "");
    while(!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
    return s;
  }
  static std::string exe_path() {
    std::vector<char> buf(PATH_MAX + 1, '\0');
    const ssize_t l = readlink("/proc/self/exe", buf.data(), buf.size() - 1);
    return l > 0 ? std::string(buf.data(), (size_t)l) : std::string();
  }
public:
  file_header() : root_(json::object()), offset_(0) { root_["alignment"] = json(8); }

  const json& root() const { return root_; }
  json& root() { return root_; }
  size_t offset() const { return offset_; }
  int alignment() const { int a = (int)root_.at("alignment").as_i64(0); return a > 0 ? a : 0; }

  // -- generic part (the keys generic_file_header.hpp:113-152 records) ----------------------
  void fill_standard() {
    const char* epoch = getenv("SOURCE_DATE_EPOCH");
    root_["hostname"] = json(host_name(epoch != nullptr));
    root_["pwd"] = json(work_dir(epoch != nullptr));
    root_["time"] = json(time_stamp(epoch));
    root_["exe_path"] = json(exe_path());
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
    const std::string framed = frame(root_.dump(), (size_t)alignment());
    os.write(framed.data(), (std::streamsize)framed.size());
    offset_ = framed.size();
  }

  // parse from memory (mmap'd database); returns false on failure
  bool read(const char* data, size_t size) {
    size_t lo = 0, hi = 0, body = 0;
    if(!unframe(data, size, &lo, &hi, &body)) return false;
    offset_ = body;
    return json::parse(data + lo, data + hi, root_);
  }
  bool read(std::istream& is) {
    std::string buf(LEN_DIGITS, '\0');
    is.read(&buf[0], (std::streamsize)LEN_DIGITS);
    if((size_t)is.gcount() != LEN_DIGITS) return false;
    size_t d = 0; unsigned long len = 0;
    while(d < LEN_DIGITS && buf[d] >= '0' && buf[d] <= '9') { len = len * 10 + (unsigned long)(buf[d] - '0'); ++d; }
    if(d != LEN_DIGITS || len < 2) return false;
    buf.resize(LEN_DIGITS + len);
    is.read(&buf[LEN_DIGITS], (std::streamsize)len);
    if((size_t)is.gcount() != len) return false;
    return read(buf.data(), buf.size());
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
