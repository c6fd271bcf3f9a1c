// Tuned compare kernels (dispatch stub: nothing specialised yet, the generic
// kernel handles every batch).
#include "common.cuh"

int chess_launch_compare_fast(chess_ctx *, const double *, const double *, const chess_pair *, int64_t,
                              int32_t, const chess_params &, double *, double *, int32_t *, cudaStream_t,
                              bool *handled) {
  *handled = false;
  return CHESS_OK;
}
