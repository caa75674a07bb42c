// Stand-in (spec: reference sub_commands/bc_main_cmdline.yaggo). ORACLE BUILD ONLY.
#ifndef __BC_MAIN_CMDLINE_HPP__
#define __BC_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class bc_main_cmdline {
public:
  uint64_t size_arg; uint32_t mer_len_arg; double fpr_arg; bool canonical_flag; uint32_t threads_arg;
  const char* output_arg; uint32_t Files_arg; bool generator_given; const char* generator_arg;
  uint32_t Generators_arg; bool shell_given; const char* shell_arg; bool timing_given; const char* timing_arg;
  std::vector<const char*> file_arg;
  bc_main_cmdline() : size_arg(0), mer_len_arg(0), fpr_arg(0.001), canonical_flag(false), threads_arg(1),
    output_arg("mer_bloom_filter"), Files_arg(1), generator_given(false), generator_arg(""), Generators_arg(1),
    shell_given(false), shell_arg(""), timing_given(false), timing_arg("") { }
  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }
  void parse(int argc, char* argv[]) {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish bc [options] file:path+");
    p.add("size", 's', U64S, &size_arg, 0, true);
    p.add("mer-len", 'm', U32, &mer_len_arg, 0, true);
    p.add("fpr", 'f', DOUBLE, &fpr_arg);
    p.add("canonical", 'C', FLAG, &canonical_flag);
    p.add("threads", 't', U32, &threads_arg);
    p.add("output", 'o', CSTR, &output_arg);
    p.add("Files", 'F', U32, &Files_arg);
    p.add("generator", 'g', CSTR, &generator_arg, &generator_given);
    p.add("Generators", 'G', U32, &Generators_arg);
    p.add("shell", 'S', CSTR, &shell_arg, &shell_given);
    p.add("timing", 0, CSTR, &timing_arg, &timing_given);
    p.parse(argc, argv, file_arg);
  }
};
#endif
