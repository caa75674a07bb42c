// Stand-in (spec: reference sub_commands/info_main_cmdline.yaggo). ORACLE BUILD ONLY.
#ifndef __INFO_MAIN_CMDLINE_HPP__
#define __INFO_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class info_main_cmdline {
public:
  bool skip_flag, json_flag, cmd_flag; const char* file_arg;
  info_main_cmdline() : skip_flag(false), json_flag(false), cmd_flag(false), file_arg("") { }
  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }
  void parse(int argc, char* argv[]) {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish info [options] file:path");
    p.add("skip", 's', FLAG, &skip_flag);
    p.add("json", 'j', FLAG, &json_flag);
    p.add("cmd", 'c', FLAG, &cmd_flag);
    std::vector<const char*> pos;
    p.parse(argc, argv, pos);
    if(pos.size() != 1) error("Requires exactly 1 argument.");
    file_arg = pos[0];
  }
};
#endif
