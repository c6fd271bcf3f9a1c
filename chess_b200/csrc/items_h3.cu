// Work-item compare kernel instantiations for the windowed regime with SSIM window 7 (-a 7),
// column-strip widths G = 3..8, single-block and (G >= 5) multi-block.  Separate translation unit so
// that the library builds in parallel.
#include "compare_kernels.cuh"

int chess_items_launch_h3(chess_ctx *ctx, chess_item_args &a, int G, bool multi, cudaStream_t stream) {
  if (multi) {  // matrices wider than one block: strips of 5..8 columns, several column blocks
    switch (G) {
      case 5: return launch_items_sn<5, true, 3, true>(ctx, a, stream);
      case 6: return launch_items_sn<6, true, 3, true>(ctx, a, stream);
      case 7: return launch_items_sn<7, true, 3, true>(ctx, a, stream);
      default: return launch_items_sn<8, true, 3, true>(ctx, a, stream);
    }
  }
  switch (G) {
    case 3: return launch_items_sn<3, true, 3, false>(ctx, a, stream);
    case 4: return launch_items_sn<4, true, 3, false>(ctx, a, stream);
    case 5: return launch_items_sn<5, true, 3, false>(ctx, a, stream);
    case 6: return launch_items_sn<6, true, 3, false>(ctx, a, stream);
    case 7: return launch_items_sn<7, true, 3, false>(ctx, a, stream);
    default: return launch_items_sn<8, true, 3, false>(ctx, a, stream);
  }
}
