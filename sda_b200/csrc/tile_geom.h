// tile_geom.h -- decomposition of the output pixels of a convolution into 128-pixel M tiles.
//
// A tile is a box of BW x BH pixels of BN consecutive images with BW * BH * BN = 128, so that the
// input patch of every 3x3 tap is ONE dense TMA box of the haloed operand tensor and the 128 rows
// of the implicit-GEMM A tile are the box in (image, row, column) order.
#pragma once

#include "common.cuh"

namespace sdab {

struct TileGeom {
  int BW, BH, BN;
  int tiles_w, tiles_h, tiles_n, num_tiles;

  __host__ __device__ __forceinline__ void tile_origin(int tile, int& n0, int& h0, int& w0) const {
    const int tw = tile % tiles_w;
    const int th = (tile / tiles_w) % tiles_h;
    const int tn = tile / (tiles_w * tiles_h);
    n0 = tn * BN, h0 = th * BH, w0 = tw * BW;
  }
};

inline int make_tile_geom(int N, int H, int W, TileGeom& g) {
  SDAB_REQUIRE(N >= 1 && H >= 1 && W >= 1, "empty convolution");
  if (W >= 128) {
    SDAB_REQUIRE(W % 128 == 0, "image width must be a power of two below 128 or a multiple of 128");
    g.BW = 128, g.BH = 1, g.BN = 1;
  } else {
    SDAB_REQUIRE(128 % W == 0, "image width must be a power of two below 128 or a multiple of 128");
    g.BW = W;
    const int rows = 128 / W;
    if (H >= rows) {
      SDAB_REQUIRE(H % rows == 0, "image height must be a multiple of 128 / width");
      g.BH = rows;
    } else {
      SDAB_REQUIRE(rows % H == 0, "image height must divide 128 / width");
      g.BH = H;
    }
    g.BN = 128 / (g.BW * g.BH);
  }
  g.tiles_w = W / g.BW;
  g.tiles_h = H / g.BH;
  g.tiles_n = (N + g.BN - 1) / g.BN;
  g.num_tiles = g.tiles_w * g.tiles_h * g.tiles_n;
  return SDAB_OK;
}

}  // namespace sdab
