// recommendK scoring (placeholder until the tcgen05 kernel lands in this file).
#include "common.cuh"
using namespace lgc;
extern "C" size_t lgc_score_topk_workspace_bytes(int64_t, int64_t, int, int) { return 0; }
extern "C" int lgc_score_topk(const lgc_score_topk_args*, void*) {
  set_error("lgc_score_topk: not built yet");
  return LGC_ERR_UNSUPPORTED;
}
