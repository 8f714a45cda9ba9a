// Host-side lowering of a GEMM group onto the tcgen05 kernel: bf16 image handles, TMA tensor maps
// (cuTensorMapEncodeTiled obtained through the runtime's driver entry point, no -lcuda), tiling.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "gemm_tc.cuh"

namespace dsact {

// bf16 hi/lo image of a row-major [rows, width] fp32 tensor
struct Img {
  __nv_bfloat16* p = nullptr;
  int rows = 0, width = 0, pitch = 0;
  long long plane = 0;
  Img cols(int c0, int w) const { Img v = *this; v.p += c0; v.width = w; return v; }  // column slice (c0 % 8 == 0)
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// dims (width, rows, 2 planes); box (64, box_rows, 1); 128-byte swizzle; out-of-bounds elements read as zero.
static bool make_map(CUtensorMap* m, const Img& t, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)t.width, (cuuint64_t)t.rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)t.pitch * 2, (cuuint64_t)t.plane * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, t.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 [rows, width] row-major tensor (leading dimension ld floats), box 16 x 32: epilogue act' tiles.  64-byte swizzle
// on the shared-memory side (16-byte chunk ^= (row >> 1) & 3): the epilogue's lane-per-row float4 accesses, 64 bytes
// apart, are then bank-conflict free (unswizzled they are 4-way conflicts).
static bool make_map_f32(CUtensorMap* m, const float* base, int rows, int width, int ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// bf16 image as a store target: dims (width, rows, 2 planes), box 16 x 32 x 1, 32-byte swizzle (16-byte chunk ^= (row >> 2) & 1)
// `planes` = 2: one request stores the hi tile and the lo tile behind it (shared memory: [hi 1 KiB][lo 1 KiB])
static bool make_map_img_store(CUtensorMap* m, const Img& t, int planes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)t.width, (cuuint64_t)t.rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)t.pitch * 2, (cuuint64_t)t.plane * 2};
  cuuint32_t box[3] = {16, 32, (cuuint32_t)planes};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, t.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// what the tcgen05 lowering needs beyond a GemmProb
struct TcExtra {
  Img a[2], b, out;
  int kB0[2] = {0, 0};
};

inline int tc_smem_bytes(int stages, int planes, int stage_b) {
  return stages * planes * (TC_STAGE_A + stage_b) + (2 * stages + 2) * 8 + 1024;
}

}  // namespace dsact
