// Instantiations of the block-triangle windowed kernels (compare_blk7.cuh, compare_blkw.cuh), G = 5..8.
#include "compare_blk7.cuh"
#include "compare_blkw.cuh"

template <int H>
static int launch_h(chess_ctx *ctx, chess_item_args &a, int G, cudaStream_t stream) {
  switch (G) {
    case 5: return launch_blkw<5, H>(ctx, a, stream);
    case 6: return launch_blkw<6, H>(ctx, a, stream);
    case 7: return launch_blkw<7, H>(ctx, a, stream);
    default: return launch_blkw<8, H>(ctx, a, stream);
  }
}

int chess_items_launch_blkw(chess_ctx *ctx, chess_item_args &a, int G, int H, cudaStream_t stream) {
  switch (H) {
    case 1: return launch_h<1>(ctx, a, G, stream);
    case 2: return launch_h<2>(ctx, a, G, stream);
    case 3:
      switch (G) {
        case 5: return launch_blk7<5>(ctx, a, stream);
        case 6: return launch_blk7<6>(ctx, a, stream);
        case 7: return launch_blk7<7>(ctx, a, stream);
        default: return launch_blk7<8>(ctx, a, stream);
      }
    case 4: return launch_h<4>(ctx, a, G, stream);
    default: return launch_h<5>(ctx, a, G, stream);
  }
}
