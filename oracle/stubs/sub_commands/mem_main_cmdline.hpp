// Stand-in (spec: reference sub_commands/mem_main_cmdline.yaggo; ignored switches omitted). ORACLE BUILD ONLY.
#ifndef __MEM_MAIN_CMDLINE_HPP__
#define __MEM_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class mem_main_cmdline {
public:
  uint32_t mer_len_arg; bool size_given; uint64_t size_arg; uint32_t counter_len_arg, reprobes_arg;
  bool mem_given; uint64_t mem_arg;
  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }
  mem_main_cmdline(int argc, char* argv[]) : mer_len_arg(0), size_given(false), size_arg(0), counter_len_arg(7),
    reprobes_arg(126), mem_given(false), mem_arg(0) {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish mem [options]");
    p.add("mer-len", 'm', U32, &mer_len_arg, 0, true);
    p.add("size", 's', U64S, &size_arg, &size_given);
    p.add("counter-len", 'c', U32, &counter_len_arg);
    p.add("reprobes", 'p', U32, &reprobes_arg);
    p.add("mem", 0, U64S, &mem_arg, &mem_given);
    std::vector<const char*> pos;
    p.parse(argc, argv, pos);
    if(size_given && mem_given) error("Switches [--mem] and [-s, --size] conflict");
  }
};
#endif
