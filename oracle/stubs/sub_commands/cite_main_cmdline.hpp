// Stand-in (spec: reference sub_commands/cite_main_cmdline.yaggo). ORACLE BUILD ONLY.
#ifndef __CITE_MAIN_CMDLINE_HPP__
#define __CITE_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class cite_main_cmdline {
public:
  bool bibtex_flag; bool output_given; const char* output_arg;
  cite_main_cmdline(int argc, char* argv[]) : bibtex_flag(false), output_given(false), output_arg("") {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish cite [options]");
    p.add("bibtex", 'b', FLAG, &bibtex_flag);
    p.add("output", 'o', CSTR, &output_arg, &output_given);
    std::vector<const char*> pos;
    p.parse(argc, argv, pos);
  }
};
#endif
