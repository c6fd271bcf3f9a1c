// Instantiations of the upper-triangle -a 3 / 5 / 9 / 11 kernel (compare_triw.cuh), strip widths G = max(3, H)..8.
#include "compare_triw.cuh"

template <int H>
static int launch_h(chess_ctx *ctx, chess_item_args &a, int G, cudaStream_t stream) {
  switch (G) {
    case 3: if constexpr (H <= 3) return launch_triw<3, H>(ctx, a, stream);
    case 4: if constexpr (H <= 4) return launch_triw<4, H>(ctx, a, stream);
    case 5: return launch_triw<5, H>(ctx, a, stream);
    case 6: return launch_triw<6, H>(ctx, a, stream);
    case 7: return launch_triw<7, H>(ctx, a, stream);
    default: return launch_triw<8, H>(ctx, a, stream);
  }
}

int chess_items_launch_triw(chess_ctx *ctx, chess_item_args &a, int G, int H, cudaStream_t stream) {
  switch (H) {
    case 1: return launch_h<1>(ctx, a, G, stream);
    case 2: return launch_h<2>(ctx, a, G, stream);
    case 4: return launch_h<4>(ctx, a, G, stream);   // G >= 4 (chess_launch_compare_items picks it)
    default: return launch_h<5>(ctx, a, G, stream);  // G >= 5
  }
}
