// Stand-in (spec: reference sub_commands/dump_main_cmdline.yaggo). ORACLE BUILD ONLY.
#ifndef __DUMP_MAIN_CMDLINE_HPP__
#define __DUMP_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class dump_main_cmdline {
public:
  bool column_flag, tab_flag; bool lower_count_given, upper_count_given;
  uint64_t lower_count_arg, upper_count_arg; bool output_given; const char* output_arg; const char* db_arg;
  dump_main_cmdline() : column_flag(false), tab_flag(false), lower_count_given(false), upper_count_given(false),
    lower_count_arg(0), upper_count_arg(0), output_given(false), output_arg(""), db_arg("") { }
  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }
  void parse(int argc, char* argv[]) {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish dump [options] db:path");
    p.add("column", 'c', FLAG, &column_flag);
    p.add("tab", 't', FLAG, &tab_flag);
    p.add("lower-count", 'L', U64, &lower_count_arg, &lower_count_given);
    p.add("upper-count", 'U', U64, &upper_count_arg, &upper_count_given);
    p.add("output", 'o', CSTR, &output_arg, &output_given);
    std::vector<const char*> pos;
    p.parse(argc, argv, pos);
    if(pos.size() != 1) error("Requires exactly 1 argument.");
    db_arg = pos[0];
  }
};
#endif
