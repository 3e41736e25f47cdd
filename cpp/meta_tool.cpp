// cuNVSMMeta — writes / prints the `_meta` file of a model (lse.Metadata, proto/nvsm.proto:91-108) without touching
// the GPU:  cuNVSMMeta write <ngram_file> <window_size> <output>   ->  <output>_meta  (what cuNVSMTrainModel writes,
//                                                                      reference: cpp/main.cu:527-537)
//           cuNVSMMeta print <meta_file>                            ->  one line per term / object (any lse.Metadata file,
//                                                                      the reference's own included)
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

#include "cuNVSM/data.h"

int main(int argc, char** argv) {
  if (argc == 5 && std::strcmp(argv[1], "write") == 0) {
    TextEntity::NGramFileSource source(argv[2], std::stoul(argv[3]), /*rng=*/nullptr, /*no_shuffle=*/true);
    lse::Metadata meta;
    source.extract_metadata(&meta);
    const std::string path = std::string(argv[4]) + "_meta";
    std::ofstream out(path, std::ios::binary);
    NVSM_CHECK(meta.SerializeToOstream(&out), "cannot write the _meta file");
    out.close();
    // read back what was written: the parser must reproduce the message
    lse::Metadata again;
    std::ifstream in(path, std::ios::binary);
    NVSM_CHECK(again.ParseFromIstream(&in), "written _meta does not parse");
    NVSM_CHECK(again.SerializeAsString() == meta.SerializeAsString(), "_meta round trip differs");
    std::printf("terms %d objects %d total_terms %d\n", meta.term_size(), meta.object_size(), meta.total_terms());
    return 0;
  }
  if (argc == 3 && std::strcmp(argv[1], "print") == 0) {
    lse::Metadata meta;
    std::ifstream in(argv[2], std::ios::binary);
    NVSM_CHECK(in.good(), "cannot open the meta file");
    NVSM_CHECK(meta.ParseFromIstream(&in), "not an lse.Metadata message");
    for (int i = 0; i < meta.term_size(); ++i)
      std::printf("term %d %d %d\n", meta.term(i).index_term_id(), meta.term(i).model_term_id(), meta.term(i).term_frequency());
    for (int i = 0; i < meta.object_size(); ++i)
      std::printf("object %d %d\n", meta.object(i).index_object_id(), meta.object(i).model_object_id());
    std::printf("total_terms %d\n", meta.total_terms());
    return 0;
  }
  std::fprintf(stderr, "usage: %s write <ngram_file> <window_size> <output> | print <meta_file>\n", argv[0]);
  return 2;
}
