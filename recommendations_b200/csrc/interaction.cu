// Ranker pairwise dot interaction on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// No reference implementation exists (models/ranker/fdlrm/ is empty): canonical DLRM
// interaction.  feats bf16 [B, F, D] -> out bf16 [B, F(F-1)/2] = strictly-lower triangle of
// feats[b] feats[b]^T with fp32 accumulation.
//
// The op is HBM-bound (AI ~ 12-34 flop/B vs. a ridge of ~210): the tensor cores are there so
// the math hides under the feature read.  A CTA (4 warps) works on tiles of 4 samples: their F
// rows are padded to 32, giving a 128-row bf16 tile that is BOTH operands of one Gram product
// D[128x128] = T T^T (K-major A and B descriptors over the same shared-memory tile, SWIZZLE_128B
// canonical layout written with cp.async), accumulated in 128 TMEM columns.  Only the four
// 32x32 diagonal blocks are read back: warp w owns TMEM lanes 32w..32w+31 = sample w, one
// tcgen05.ld.32x32b.x32 gives lane i the row Z[i, 0..31], the j < i entries are packed through
// shared memory so the tile's 4 x F(F-1)/2 outputs leave as one contiguous coalesced span.
// Several CTAs per SM (32 KB smem, 128 TMEM columns each) overlap load / MMA / epilogue.
//
// Backward: grad_feats[b] = (G + G^T) feats[b] with G the lower-triangular unpack of
// grad_out[b].  Same tile, now the MN-major B operand; A = blockdiag(G_s + G_s^T) [128x128] is
// built in shared memory per tile; D[128 x D] in TMEM, staged through smem, written coalesced.
#include "common.cuh"

namespace recemb {

constexpr int kIxThreads = 128;
constexpr int kIxSamples = 4;    // samples per tile
constexpr int kIxRows = 32;      // feature rows per sample after padding
constexpr int kKBlockBytes = 128 * 128;  // one 64-wide K block of the 128-row tile (SW128)

// byte offset of 16-byte chunk `c16` (0..7) of row `r` inside a [rows x 128 B] SWIZZLE_128B block
__device__ __forceinline__ uint32_t sw128(int r, int c16) { return r * 128 + ((c16 ^ (r & 7)) << 4); }

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // LayoutType::SWIZZLE_128B
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> fp32, M x N
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n, int b_mn_major) {
  return (1u << 4) /*C fp32*/ | (1u << 7) /*A bf16*/ | (1u << 10) /*B bf16*/ |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// lane i of the warp receives 32 consecutive fp32 columns of its TMEM lane
__device__ __forceinline__ void tmem_ld_32cols(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// bounded mbarrier wait: a descriptor mistake must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// cooperative load of the 4-sample tile into the SW128 K-major layout (zero rows for padding)
template <int DIM>
__device__ __forceinline__ void load_feature_tile(uint8_t* s_tile, const __nv_bfloat16* feats, int64_t tile,
                                                  int64_t batch, int F) {
  constexpr int CH = DIM / 8;  // 16-byte chunks per row
  const uint32_t s_base = smem_u32(s_tile);
  for (int idx = threadIdx.x; idx < 128 * CH; idx += kIxThreads) {
    const int r = idx / CH, c = idx % CH;
    const int s = r / kIxRows, i = r % kIxRows;
    const int64_t b = tile * kIxSamples + s;
    const uint32_t off = (c / 8) * kKBlockBytes + sw128(r, c % 8);
    if (b < batch && i < F) {
      cp_async16(s_base + off, feats + (b * F + i) * DIM + c * 8);
    } else {
      *reinterpret_cast<uint4*>(s_tile + off) = make_uint4(0, 0, 0, 0);
    }
  }
}

template <int DIM>
__global__ void __launch_bounds__(kIxThreads) dot_fwd_kernel(const __nv_bfloat16* __restrict__ feats,
                                                            __nv_bfloat16* __restrict__ out, int64_t batch,
                                                            int F) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int KB = DIM / 64;
  uint8_t* s_tile = smem;                                                      // KB x 16 KB
  __nv_bfloat16* s_out = reinterpret_cast<__nv_bfloat16*>(smem + KB * kKBlockBytes);  // 4 x P
  const int P = F * (F - 1) / 2;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + KB * kKBlockBytes + ((kIxSamples * P * 2 + 15) & ~15));
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(s_tmem, 128);
  if (threadIdx.x == 0) {
    mbar_init(s_bar, 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  constexpr uint32_t idesc = umma_idesc(128, 128, 0);
  const int64_t num_tiles = (batch + kIxSamples - 1) / kIxSamples;
  uint32_t phase = 0;

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    load_feature_tile<DIM>(s_tile, feats, tile, batch, F);
    cp_async_wait_all();
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t s_base = smem_u32(s_tile);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {  // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle row
          const uint64_t d = umma_desc(s_base + kb * kKBlockBytes + k4 * 32, 16, 1024);
          umma_bf16(tmem, d, d, idesc, (kb | k4) ? 1u : 0u);
        }
      }
      umma_commit(s_bar);  // arrives when the MMAs (and their smem reads) are done
    }
    mbar_wait_bounded(s_bar, phase);
    phase ^= 1;
    tc_fence_after();

    // warp w = sample w: rows 32w.. of D, columns 32w..32w+31 (its diagonal block)
    float v[32];
    tmem_ld_32cols(tmem + ((uint32_t)(warp * 32) << 16) + warp * 32, v);
    const int64_t b = tile * kIxSamples + warp;
    if (b < batch && lane < F) {
      __nv_bfloat16* dst = s_out + warp * P + lane * (lane - 1) / 2;
#pragma unroll
      for (int j = 0; j < 31; ++j)
        if (j < lane) dst[j] = __float2bfloat16_rn(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    // the tile's valid outputs are one contiguous span of `cnt` bf16
    const int nsamp = (int)min((int64_t)kIxSamples, batch - tile * kIxSamples);
    const int cnt = nsamp * P;
    __nv_bfloat16* gout = out + tile * kIxSamples * P;
    if ((((uintptr_t)gout) & 3) == 0) {
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s_out);
      uint32_t* g32 = reinterpret_cast<uint32_t*>(gout);
      for (int i = threadIdx.x; i < cnt / 2; i += kIxThreads) g32[i] = s32[i];
      if ((cnt & 1) && threadIdx.x == 0) gout[cnt - 1] = s_out[cnt - 1];
    } else {
      for (int i = threadIdx.x; i < cnt; i += kIxThreads) gout[i] = s_out[i];
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

template <int DIM>
__global__ void __launch_bounds__(kIxThreads) dot_bwd_kernel(const __nv_bfloat16* __restrict__ feats,
                                                            const __nv_bfloat16* __restrict__ grad_out,
                                                            __nv_bfloat16* __restrict__ grad_feats,
                                                            int64_t batch, int F) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NB = DIM / 64;            // MN atoms (64 dims each) of the B operand
  constexpr int STG = DIM * 2;            // staging row (bytes); 16-byte chunks XOR-swizzled by the row
  constexpr int CHS = DIM / 8;            // 16-byte chunks per staging row (8 or 16: a power of two)
  uint8_t* s_tile = smem;                 // B operand: NB x 16 KB (rows = K index, 128 B = 64 dims)
  uint8_t* s_a = smem + NB * kKBlockBytes;            // A = blockdiag(G+G^T): 2 K-blocks x 16 KB
  // the output staging (128 x STG = the size of the B operand) reuses the B operand's memory: it is
  // written after the MMAs have committed (B is dead) and the next tile reloads B anyway.  64 KB per
  // CTA instead of 99 KB: 3 CTAs (12 warps, 3 x 128 TMEM columns) per SM overlap load / MMA / epilogue
  uint8_t* s_stage = s_tile;
  uint16_t* s_ij = reinterpret_cast<uint16_t*>(s_a + 2 * kKBlockBytes);   // e -> (i << 8 | j), P entries
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_a + 2 * kKBlockBytes + 1024);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);
  constexpr uint32_t TMEM_COLS = DIM <= 64 ? 64 : (DIM <= 128 ? 128 : 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = F * (F - 1) / 2;
  if (warp == 0) tmem_alloc(s_tmem, TMEM_COLS);
  if (threadIdx.x == 0) {
    mbar_init(s_bar, 1);
    mbar_fence_init();
  }
  // the off-diagonal blocks of A are zero for every tile: clear once
  for (int i = threadIdx.x; i < 2 * kKBlockBytes / 16; i += kIxThreads)
    reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0, 0, 0, 0);
  // triangle index e = i(i-1)/2 + j -> (i, j), once per CTA instead of a sqrt per element per tile
  for (int e = threadIdx.x; e < P; e += kIxThreads) {
    int i = (int)((1.f + sqrtf(1.f + 8.f * (float)e)) * 0.5f);
    while (i * (i - 1) / 2 > e) --i;
    while ((i + 1) * i / 2 <= e) ++i;
    s_ij[e] = (uint16_t)((i << 8) | (e - i * (i - 1) / 2));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  constexpr uint32_t idesc = umma_idesc(128, DIM, 1);
  const int64_t num_tiles = (batch + kIxSamples - 1) / kIxSamples;
  uint32_t phase = 0;

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    load_feature_tile<DIM>(s_tile, feats, tile, batch, F);
    // A[(s,i)][(s,j)] = A[(s,j)][(s,i)] = grad_out[b, i(i-1)/2 + j]
    for (int idx = threadIdx.x; idx < kIxSamples * P; idx += kIxThreads) {
      const int s = idx / P, e = idx - s * P;
      const int ij = s_ij[e], i = ij >> 8, j = ij & 0xff;
      const int64_t b = tile * kIxSamples + s;
      __nv_bfloat16 g = __float2bfloat16_rn(0.f);
      if (b < batch) g = grad_out[b * P + e];
      const int r1 = s * kIxRows + i, k1 = s * kIxRows + j;
      *reinterpret_cast<__nv_bfloat16*>(s_a + (k1 / 64) * kKBlockBytes + sw128(r1, (k1 % 64) / 8) + (k1 % 8) * 2) = g;
      *reinterpret_cast<__nv_bfloat16*>(s_a + (r1 / 64) * kKBlockBytes + sw128(k1, (r1 % 64) / 8) + (r1 % 8) * 2) = g;
    }
    cp_async_wait_all();
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a_base = smem_u32(s_a), b_base = smem_u32(s_tile);
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16) {  // K = 128 rows of the tile, 16 per MMA
        const uint64_t da = umma_desc(a_base + (k16 / 4) * kKBlockBytes + (k16 % 4) * 32, 16, 1024);
        // MN-major B: 128-byte rows are 64 contiguous dims; LBO = next 64 dims, SBO = next 8 K rows
        const uint64_t db = umma_desc(b_base + k16 * 16 * 128, kKBlockBytes, 1024);
        umma_bf16(tmem, da, db, idesc, k16 ? 1u : 0u);
      }
      umma_commit(s_bar);
    }
    mbar_wait_bounded(s_bar, phase);
    phase ^= 1;
    tc_fence_after();

    // lane i of warp w holds row (sample w, feature i): DIM fp32 columns -> bf16 staging row
    // (chunk c of row r lives at chunk position c ^ (r % CHS): the 32 lanes of a warp write 32
    // different rows at the same logical chunk -> spread over all chunk positions, no bank pile-up)
    const int srow_i = warp * 32 + lane;
    uint8_t* srow = s_stage + srow_i * STG;
#pragma unroll
    for (int c0 = 0; c0 < DIM; c0 += 32) {
      float v[32];
      tmem_ld_32cols(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          __nv_bfloat162 p = __floats2bfloat162_rn(v[q * 8 + 2 * t], v[q * 8 + 2 * t + 1]);
          w[t] = *reinterpret_cast<uint32_t*>(&p);
        }
        *reinterpret_cast<uint4*>(srow + ((((c0 >> 3) + q) ^ (srow_i & (CHS - 1))) << 4)) =
            make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    tc_fence_before();
    __syncthreads();
    constexpr int CH = DIM / 8;
    const int nsamp = (int)min((int64_t)kIxSamples, batch - tile * kIxSamples);
    for (int idx = threadIdx.x; idx < nsamp * F * CH; idx += kIxThreads) {
      const int c = idx % CH, rowi = idx / CH;
      const int s = rowi / F, i = rowi % F;
      const int sr = s * 32 + i;
      const uint4 val = *reinterpret_cast<const uint4*>(s_stage + sr * STG + ((c ^ (sr & (CHS - 1))) << 4));
      stg_cs_v4(grad_feats + ((tile * kIxSamples + s) * F + i) * DIM + c * 8, val);
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

static size_t fwd_smem(int dim, int F) {
  const int P = F * (F - 1) / 2;
  return (size_t)(dim / 64) * kKBlockBytes + ((kIxSamples * P * 2 + 15) & ~15) + 32;
}
static size_t bwd_smem(int dim) {
  // B operand (also the output staging) + A + (i, j) table (1 KB) + barrier / TMEM slot
  return (size_t)(dim / 64) * kKBlockBytes + 2 * kKBlockBytes + 1024 + 32;
}

}  // namespace recemb

using namespace recemb;

extern "C" int recemb_dot_interaction_fwd(const void* feats, int64_t batch, int32_t num_feats, int32_t dim,
                                          void* out, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(batch >= 0, "batch < 0");
  RECEMB_UNSUPPORTED(num_feats >= 2 && num_feats <= kIxRows, "num_feats %d outside [2, 32]", num_feats);
  RECEMB_UNSUPPORTED(dim == 64 || dim == 128 || dim == 256, "dim %d not in {64, 128, 256}", dim);
  if (batch == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(feats && out, "null pointer");
  RECEMB_CHECK_ARG((uintptr_t)feats % 16 == 0, "feats must be 16-byte aligned");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const size_t smem = fwd_smem(dim, num_feats);
  const int64_t tiles = (batch + kIxSamples - 1) / kIxSamples;
  const int ctas_per_sm = dim == 256 ? 2 : 4;  // 4 x 128 TMEM columns fill the 512 of an SM
  int64_t grid = (int64_t)sm_count(device) * ctas_per_sm;
  if (grid > tiles) grid = tiles;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_FWD(D_)                                                                                  \
  {                                                                                                     \
    RECEMB_CUDA(cudaFuncSetAttribute(dot_fwd_kernel<D_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                     (int)smem));                                                       \
    dot_fwd_kernel<D_><<<(unsigned)grid, kIxThreads, smem, s>>>(                                         \
        (const __nv_bfloat16*)feats, (__nv_bfloat16*)out, batch, num_feats);                            \
  }
  if (dim == 64) LAUNCH_FWD(64) else if (dim == 128) LAUNCH_FWD(128) else LAUNCH_FWD(256)
#undef LAUNCH_FWD
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

extern "C" int recemb_dot_interaction_bwd(const void* feats, const void* grad_out, int64_t batch,
                                          int32_t num_feats, int32_t dim, void* grad_feats, int device,
                                          recemb_stream_t stream) {
  RECEMB_CHECK_ARG(batch >= 0, "batch < 0");
  RECEMB_UNSUPPORTED(num_feats >= 2 && num_feats <= kIxRows, "num_feats %d outside [2, 32]", num_feats);
  RECEMB_UNSUPPORTED(dim == 64 || dim == 128, "dim %d not in {64, 128}", dim);
  if (batch == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(feats && grad_out && grad_feats, "null pointer");
  RECEMB_CHECK_ARG(((uintptr_t)feats | (uintptr_t)grad_feats) % 16 == 0, "feats / grad_feats must be 16-byte aligned");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const size_t smem = bwd_smem(dim);
  const int64_t tiles = (batch + kIxSamples - 1) / kIxSamples;
  int64_t grid = (int64_t)sm_count(device) * 3;
  if (grid > tiles) grid = tiles;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_BWD(D_)                                                                                  \
  {                                                                                                     \
    RECEMB_CUDA(cudaFuncSetAttribute(dot_bwd_kernel<D_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                     (int)smem));                                                       \
    dot_bwd_kernel<D_><<<(unsigned)grid, kIxThreads, smem, s>>>(                                         \
        (const __nv_bfloat16*)feats, (const __nv_bfloat16*)grad_out, (__nv_bfloat16*)grad_feats, batch,  \
        num_feats);                                                                                     \
  }
  if (dim == 64) LAUNCH_BWD(64) else LAUNCH_BWD(128)
#undef LAUNCH_BWD
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}
