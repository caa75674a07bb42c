// Stand-in (spec: reference sub_commands/histo_main_cmdline.yaggo). ORACLE BUILD ONLY.
#ifndef __HISTO_MAIN_CMDLINE_HPP__
#define __HISTO_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class histo_main_cmdline {
public:
  uint64_t low_arg, high_arg, increment_arg; uint32_t threads_arg; bool full_flag;
  bool output_given; const char* output_arg; uint64_t buffer_size_arg; bool verbose_flag; const char* db_arg;
  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }
  histo_main_cmdline(int argc, char* argv[]) : low_arg(1), high_arg(10000), increment_arg(1), threads_arg(1),
    full_flag(false), output_given(false), output_arg(""), buffer_size_arg(10000000), verbose_flag(false), db_arg("") {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish histo [options] db:path");
    p.add("low", 'l', U64, &low_arg);
    p.add("high", 'h', U64, &high_arg);
    p.add("increment", 'i', U64, &increment_arg);
    p.add("threads", 't', U32, &threads_arg);
    p.add("full", 'f', FLAG, &full_flag);
    p.add("output", 'o', CSTR, &output_arg, &output_given);
    p.add("buffer-size", 's', U64S, &buffer_size_arg);
    p.add("verbose", 'v', FLAG, &verbose_flag);
    std::vector<const char*> pos;
    p.parse(argc, argv, pos);
    if(pos.size() != 1) error("Requires exactly 1 argument.");
    db_arg = pos[0];
  }
};
#endif
