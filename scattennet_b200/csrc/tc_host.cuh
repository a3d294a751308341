// Host-side tensor-map encoders shared by the tcgen05 kernels (defined in gemm_tc.cu).
#pragma once

#include <cuda.h>

#include <cstdint>

namespace scatt {

// split planes [2][rows][K] -> boxes of 64 K-elements x box_rows rows x box_planes planes, 128-byte swizzle
// (the K-major operand tiles of the UMMA descriptors; with box_planes = 2 one TMA operation fetches the hi tile
// and, box_rows * 128 bytes behind it, the lo tile)
int encode_planes_map(CUtensorMap* map, const void* planes, int64_t rows, int K, int box_rows, int fmt, int box_planes = 1);

// output maps written by one epilogue warp at a time: fp32 [M][ldy] as 32 x 32 boxes (128-byte swizzle) and / or
// split planes [2][M][N] as 32 x 32 x 1 boxes (64-byte swizzle); either pointer may be null
int encode_out_maps(CUtensorMap* map_y, CUtensorMap* map_p, float* y, int64_t ldy, void* planes, int64_t M, int N, int fmt);

}  // namespace scatt
