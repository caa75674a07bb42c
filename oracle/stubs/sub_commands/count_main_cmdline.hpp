// Stand-in for the yaggo-generated header (spec: reference sub_commands/count_main_cmdline.yaggo:4-112).
// ORACLE BUILD ONLY.
#ifndef __COUNT_MAIN_CMDLINE_HPP__
#define __COUNT_MAIN_CMDLINE_HPP__
#include <yaggo_lite.hpp>
class count_main_cmdline {
public:
  uint32_t mer_len_arg; uint64_t size_arg; uint32_t threads_arg;
  bool sam_given; std::vector<const char*> sam_arg;
  uint32_t Files_arg; bool generator_given; const char* generator_arg; uint32_t Generators_arg;
  bool shell_given; const char* shell_arg; const char* output_arg;
  uint32_t counter_len_arg; uint32_t out_counter_len_arg; bool canonical_flag;
  bool bc_given; const char* bc_arg; bool bf_size_given; uint64_t bf_size_arg; double bf_fp_arg;
  bool if_given; std::vector<const char*> if_arg;
  bool min_qual_char_given; std::string min_qual_char_arg;
  int32_t quality_start_arg; bool quality_start_given; bool min_quality_given; int32_t min_quality_arg;
  uint32_t reprobes_arg; bool text_flag, disk_flag, no_merge_flag, no_unlink_flag;
  bool lower_count_given, upper_count_given; uint64_t lower_count_arg, upper_count_arg;
  bool timing_given; const char* timing_arg; bool no_write_flag;
  std::vector<const char*> file_arg;

  count_main_cmdline() :
    mer_len_arg(0), size_arg(0), threads_arg(1), sam_given(false), Files_arg(1),
    generator_given(false), generator_arg(""), Generators_arg(1), shell_given(false), shell_arg(""),
    output_arg("mer_counts.jf"), counter_len_arg(7), out_counter_len_arg(4), canonical_flag(false),
    bc_given(false), bc_arg(""), bf_size_given(false), bf_size_arg(0), bf_fp_arg(0.01),
    if_given(false), min_qual_char_given(false), quality_start_arg(64), quality_start_given(false),
    min_quality_given(false), min_quality_arg(0), reprobes_arg(126),
    text_flag(false), disk_flag(false), no_merge_flag(false), no_unlink_flag(false),
    lower_count_given(false), upper_count_given(false), lower_count_arg(0), upper_count_arg(0),
    timing_given(false), timing_arg(""), no_write_flag(false) { }

  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }

  void parse(int argc, char* argv[]) {
    using namespace yaggo_lite;
    parser p("Usage: jellyfish count [options] file:path+");
    p.add("mer-len", 'm', U32, &mer_len_arg, 0, true);
    p.add("size", 's', U64S, &size_arg, 0, true);
    p.add("threads", 't', U32, &threads_arg);
    p.add("sam", 0, CSTR_M, &sam_arg, &sam_given);
    p.add("Files", 'F', U32, &Files_arg);
    p.add("generator", 'g', CSTR, &generator_arg, &generator_given);
    p.add("Generators", 'G', U32, &Generators_arg);
    p.add("shell", 'S', CSTR, &shell_arg, &shell_given);
    p.add("output", 'o', CSTR, &output_arg);
    p.add("counter-len", 'c', U32, &counter_len_arg);
    p.add("out-counter-len", 0, U32, &out_counter_len_arg);
    p.add("canonical", 'C', FLAG, &canonical_flag);
    p.add("bc", 0, CSTR, &bc_arg, &bc_given);
    p.add("bf-size", 0, U64S, &bf_size_arg, &bf_size_given);
    p.add("bf-fp", 0, DOUBLE, &bf_fp_arg);
    p.add("if", 0, CSTR_M, &if_arg, &if_given);
    p.add("min-qual-char", 'Q', STRING, &min_qual_char_arg, &min_qual_char_given);
    p.add("quality-start", 0, I32, &quality_start_arg, &quality_start_given);
    p.add("min-quality", 0, I32, &min_quality_arg, &min_quality_given);
    p.add("reprobes", 'p', U32, &reprobes_arg);
    p.add("text", 0, FLAG, &text_flag);
    p.add("disk", 0, FLAG, &disk_flag);
    p.add("no-merge", 0, FLAG, &no_merge_flag);
    p.add("no-unlink", 0, FLAG, &no_unlink_flag);
    p.add("lower-count", 'L', U64, &lower_count_arg, &lower_count_given);
    p.add("upper-count", 'U', U64, &upper_count_arg, &upper_count_given);
    p.add("timing", 0, CSTR, &timing_arg, &timing_given);
    p.add("no-write", 0, FLAG, &no_write_flag);
    p.parse(argc, argv, file_arg);
    if(bc_given && bf_size_given) error("Switches [--bf-size] and [--bc] conflict");
    if(min_qual_char_given && (quality_start_given || min_quality_given))
      error("Switches [--quality-start]/[--min-quality] and [-Q, --min-qual-char] conflict");
  }
};
#endif
