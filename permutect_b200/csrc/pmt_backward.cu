// Backward half of the C-ABI (placeholder until the backward kernels land).
#include "pmt_host.h"

size_t pmt_backward_workspace_bytes(const pmt::Plan&, const PmtBatch*) { return 0; }

extern "C" int pmt_backward(const PmtModelDesc*, const float*, const PmtBatch*, const PmtOutGrads*, float*, void*, size_t, void*) {
  pmt_set_error("pmt_backward: not built yet");
  return 1;
}
