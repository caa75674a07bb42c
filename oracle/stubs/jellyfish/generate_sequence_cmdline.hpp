// Stand-in (spec: reference jellyfish/generate_sequence_cmdline.yaggo). ORACLE BUILD ONLY.
#ifndef __GENERATE_SEQUENCE_ARGS_HPP__
#define __GENERATE_SEQUENCE_ARGS_HPP__
#include <yaggo_lite.hpp>
class generate_sequence_args {
public:
  long seed_arg; std::vector<uint32_t> mer_arg; const char* output_arg; bool fastq_flag;
  bool read_length_given; uint32_t read_length_arg; bool verbose_flag; std::vector<uint64_t> length_arg;
  generate_sequence_args() : seed_arg(0), output_arg("output"), fastq_flag(false), read_length_given(false),
    read_length_arg(0), verbose_flag(false) { }
  static yaggo_lite::error_stream error() { return yaggo_lite::error_stream(); }
  static yaggo_lite::error_stream error(const char* msg) { return yaggo_lite::error_stream(msg); }
  void parse(int argc, char* argv[]) {
    using namespace yaggo_lite;
    parser p("Usage: generate_sequence [options] length:uint64+");
    p.add("seed", 's', LONG, &seed_arg, 0, true);
    p.add("mer", 'm', U32_M, &mer_arg);
    p.add("output", 'o', CSTR, &output_arg);
    p.add("fastq", 'q', FLAG, &fastq_flag);
    p.add("read-length", 'r', U32, &read_length_arg, &read_length_given);
    p.add("verbose", 'v', FLAG, &verbose_flag);
    std::vector<const char*> pos;
    p.parse(argc, argv, pos);
    if(pos.empty()) error("Requires at least 1 argument.");
    for(size_t i = 0; i < pos.size(); ++i) length_arg.push_back(to_u64(pos[i], false, "length"));
  }
};
#endif
