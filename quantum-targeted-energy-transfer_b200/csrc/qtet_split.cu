// placeholder: replaced by the real split path
#include "qtet_common.cuh"
namespace qtet {
bool split_eligible(const ProblemView&) { return false; }
int launch_split(const ProblemView&, SplitWorkspace**, const double*, int64_t, int, double*, double*, int32_t*, double*, cudaStream_t) {
    set_error("split path not built"); return QTET_EUNSUPPORTED; }
void split_workspace_free(SplitWorkspace*) {}
}
