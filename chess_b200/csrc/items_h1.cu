// Work-item compare kernel instantiations for the windowed regime with SSIM window 3 (-a 3),
// column-strip widths G = 3..8.  Separate translation unit so that the library builds in parallel.
#include "compare_kernels.cuh"

int chess_items_launch_h1(chess_ctx *ctx, chess_item_args &a, int G, cudaStream_t stream) {
  switch (G) {
    case 3: return launch_items_sn<3, true, 1>(ctx, a, stream);
    case 4: return launch_items_sn<4, true, 1>(ctx, a, stream);
    case 5: return launch_items_sn<5, true, 1>(ctx, a, stream);
    case 6: return launch_items_sn<6, true, 1>(ctx, a, stream);
    case 7: return launch_items_sn<7, true, 1>(ctx, a, stream);
    default: return launch_items_sn<8, true, 1>(ctx, a, stream);
  }
}
