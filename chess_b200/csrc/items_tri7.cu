// Instantiations of the upper-triangle -a 7 kernel (compare_tri7.cuh), strip widths G = 3..8.
#include "compare_tri7.cuh"

int chess_items_launch_tri7(chess_ctx *ctx, chess_item_args &a, int G, cudaStream_t stream) {
  switch (G) {
    case 3: return launch_tri7<3>(ctx, a, stream);
    case 4: return launch_tri7<4>(ctx, a, stream);
    case 5: return launch_tri7<5>(ctx, a, stream);
    case 6: return launch_tri7<6>(ctx, a, stream);
    case 7: return launch_tri7<7>(ctx, a, stream);
    default: return launch_tri7<8>(ctx, a, stream);
  }
}
