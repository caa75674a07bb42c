// Stand-in (spec: reference sub_commands/query_main_cmdline.yaggo). ORACLE BUILD ONLY.
#ifndef __QUERY_MAIN_CMDLINE_HPP__
#define __QUERY_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class query_main_cmdline {
public:
  bool sequence_given; std::vector<const char*> sequence_arg; bool output_given; const char* output_arg;
  bool interactive_flag, load_flag, no_load_flag; const char* file_arg; std::vector<const char*> mers_arg;
  query_main_cmdline() : sequence_given(false), output_given(false), output_arg(""), interactive_flag(false),
    load_flag(false), no_load_flag(false), file_arg("") { }
  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }
  void parse(int argc, char* argv[]) {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish query [options] file:path mers:string*");
    p.add("sequence", 's', CSTR_M, &sequence_arg, &sequence_given);
    p.add("output", 'o', CSTR, &output_arg, &output_given);
    p.add("interactive", 'i', FLAG, &interactive_flag);
    p.add("load", 'l', FLAG, &load_flag);
    p.add("no-load", 'L', FLAG, &no_load_flag);
    std::vector<const char*> pos;
    p.parse(argc, argv, pos);
    if(pos.size() < 1) error("Requires at least 1 argument.");
    file_arg = pos[0];
    mers_arg.assign(pos.begin() + 1, pos.end());
  }
};
#endif
