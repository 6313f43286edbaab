// gem_b200/csrc/n2v.cu -- node2vec on the GPU (placeholder: filled in by the next milestone).
#include "common.cuh"
using namespace gemb;
extern "C" {
int gemb_n2v_alias(gemb_graph *, const double *, int32_t *, double *) {
    set_error("gemb_n2v_alias: not implemented yet");
    return GEMB_ERR_UNSUPPORTED;
}
int gemb_n2v_walks(gemb_graph *, const double *, const int32_t *, int64_t, int, int, double, double, int32_t,
                   int64_t, int64_t, int32_t *, gemb_n2v_stats *) {
    set_error("gemb_n2v_walks: not implemented yet");
    return GEMB_ERR_UNSUPPORTED;
}
int gemb_node2vec(gemb_graph *, const double *, const int32_t *, int64_t, int, int, int, int, int, double, double,
                  int32_t, int, int64_t, float *, gemb_n2v_stats *) {
    set_error("gemb_node2vec: not implemented yet");
    return GEMB_ERR_UNSUPPORTED;
}
}
