// Stand-in (spec: reference sub_commands/stats_main_cmdline.yaggo). ORACLE BUILD ONLY.
#ifndef __STATS_MAIN_CMDLINE_HPP__
#define __STATS_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class stats_main_cmdline {
public:
  bool recompute_flag; uint64_t lower_count_arg; bool upper_count_given; uint64_t upper_count_arg;
  bool verbose_flag; bool output_given; const char* output_arg; const char* db_arg;
  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }
  stats_main_cmdline(int argc, char* argv[]) : recompute_flag(false), lower_count_arg(0), upper_count_given(false),
    upper_count_arg(0), verbose_flag(false), output_given(false), output_arg(""), db_arg("") {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish stats [options] db:path");
    p.add("recompute", 'r', FLAG, &recompute_flag);
    p.add("lower-count", 'L', U64, &lower_count_arg);
    p.add("upper-count", 'U', U64, &upper_count_arg, &upper_count_given);
    p.add("verbose", 'v', FLAG, &verbose_flag);
    p.add("output", 'o', CSTR, &output_arg, &output_given);
    std::vector<const char*> pos;
    p.parse(argc, argv, pos);
    if(pos.size() != 1) error("Requires exactly 1 argument.");
    db_arg = pos[0];
  }
};
#endif
