// Instantiations of the block-triangle -r 1 kernel (compare_blk.cuh), strip widths G = 5..8.
#include "compare_blk.cuh"

int chess_items_launch_blk(chess_ctx *ctx, chess_item_args &a, int G, cudaStream_t stream) {
  if (a.p.compute_sn) {
    switch (G) {
      case 5: return launch_blk<5, true>(ctx, a, stream);
      case 6: return launch_blk<6, true>(ctx, a, stream);
      case 7: return launch_blk<7, true>(ctx, a, stream);
      default: return launch_blk<8, true>(ctx, a, stream);
    }
  }
  switch (G) {
    case 5: return launch_blk<5, false>(ctx, a, stream);
    case 6: return launch_blk<6, false>(ctx, a, stream);
    case 7: return launch_blk<7, false>(ctx, a, stream);
    default: return launch_blk<8, false>(ctx, a, stream);
  }
}
