// Instantiations of the upper-triangle -r 1 kernel (compare_tri.cuh), strip widths G = 3..8.
#include "compare_tri.cuh"

int chess_items_launch_tri(chess_ctx *ctx, chess_item_args &a, int G, cudaStream_t stream) {
  if (a.p.compute_sn) {
    switch (G) {
      case 3: return launch_tri<3, true>(ctx, a, stream);
      case 4: return launch_tri<4, true>(ctx, a, stream);
      case 5: return launch_tri_half<5, true>(ctx, a, stream);
      case 6: return launch_tri_half<6, true>(ctx, a, stream);
      case 7: return launch_tri_half<7, true>(ctx, a, stream);
      default: return launch_tri_half<8, true>(ctx, a, stream);
    }
  }
  switch (G) {
    case 3: return launch_tri<3, false>(ctx, a, stream);
    case 4: return launch_tri<4, false>(ctx, a, stream);
    case 5: return launch_tri_half<5, false>(ctx, a, stream);
    case 6: return launch_tri_half<6, false>(ctx, a, stream);
    case 7: return launch_tri_half<7, false>(ctx, a, stream);
    default: return launch_tri_half<8, false>(ctx, a, stream);
  }
}
