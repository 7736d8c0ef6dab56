// Requester-side reduction of the row-wise sharded pooled lookup: the W owners' partial pools
// of my bags (received by all-to-all) are summed in fixed owner order 0..W-1 with fp32
// accumulation (deterministic).
#include "common.cuh"

namespace recemb {

template <typename T>
__global__ void __launch_bounds__(256) sum_partials_kernel(const uint4* __restrict__ parts, int32_t world,
                                                          int64_t vecs_per_part, const float* __restrict__ row_scale,
                                                          int32_t vecs_per_row, uint4* __restrict__ out) {
  constexpr int E = Vec16<T>::kElems;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (; i < vecs_per_part; i += stride) {
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    for (int s = 0; s < world; ++s) {
      float f[E];
      Vec16<T>::unpack(ldg_nc_v4(parts + (int64_t)s * vecs_per_part + i), f);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] += f[e];
    }
    if (row_scale) {
      const float sc = row_scale[i / vecs_per_row];
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] *= sc;
    }
    stg_cs_v4(out + i, Vec16<T>::pack(acc));
  }
}

}  // namespace recemb

using namespace recemb;

extern "C" int recemb_sum_partials(const void* parts, int32_t world, int64_t rows, int32_t dim, int dtype,
                                   const float* row_scale, void* out, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(world >= 1 && rows >= 0 && dim > 0, "bad shape");
  RECEMB_CHECK_ARG(dtype == RECEMB_F32 || dtype == RECEMB_BF16, "bad dtype");
  const int64_t row_bytes = (int64_t)dim * (dtype == RECEMB_F32 ? 4 : 2);
  RECEMB_UNSUPPORTED(row_bytes % 16 == 0, "row bytes not a multiple of 16");
  if (rows == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(parts && out, "null pointer");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const int64_t vecs = rows * (row_bytes / 16);
  int64_t grid = (vecs + 255) / 256;
  const int64_t cap = (int64_t)sm_count(device) * 8;
  if (grid > cap) grid = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == RECEMB_F32)
    sum_partials_kernel<float><<<(unsigned)grid, 256, 0, s>>>((const uint4*)parts, world, vecs, row_scale,
                                                            (int32_t)(row_bytes / 16), (uint4*)out);
  else
    sum_partials_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, s>>>(
        (const uint4*)parts, world, vecs, row_scale, (int32_t)(row_bytes / 16), (uint4*)out);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}
