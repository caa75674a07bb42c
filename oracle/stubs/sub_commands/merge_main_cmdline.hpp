// Stand-in (spec: reference sub_commands/merge_main_cmdline.yaggo). ORACLE BUILD ONLY.
#ifndef __MERGE_MAIN_CMDLINE_HPP__
#define __MERGE_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class merge_main_cmdline {
public:
  const char* output_arg; bool min_flag, max_flag, jaccard_flag;
  bool lower_count_given, upper_count_given; uint64_t lower_count_arg, upper_count_arg;
  std::vector<const char*> input_arg;
  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }
  merge_main_cmdline(int argc, char* argv[]) : output_arg("mer_counts_merged.jf"), min_flag(false), max_flag(false),
    jaccard_flag(false), lower_count_given(false), upper_count_given(false), lower_count_arg(0), upper_count_arg(0) {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish merge [options] input:string+");
    p.add("output", 'o', CSTR, &output_arg);
    p.add("min", 'm', FLAG, &min_flag);
    p.add("max", 'M', FLAG, &max_flag);
    p.add("jaccard", 'j', FLAG, &jaccard_flag);
    p.add("lower-count", 'L', U64, &lower_count_arg, &lower_count_given);
    p.add("upper-count", 'U', U64, &upper_count_arg, &upper_count_given);
    p.parse(argc, argv, input_arg);
    if(min_flag && max_flag) error("Switches [-M, --max] and [-m, --min] conflict");
    if(input_arg.size() < 2) error("Requires at least 2 arguments.");
  }
};
#endif
