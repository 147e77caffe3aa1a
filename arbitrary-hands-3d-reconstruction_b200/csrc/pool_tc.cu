// Attention pooling (Hadamard_product, /root/reference/acr/model.py:103-113, with part_attention = nearest-1/2 of the
// segmentation logits minus the background channel, :126-128) as a tcgen05 GEMM fed by TMA (sm_100a).
//
// Per image the contraction over pixels is  acc[c][j] = sum_p f[p][c] * w[j][p]  with w = exp(l - m): a
// [M = 256 channels] x [N = 32 parts] x [K = pixels] GEMM whose A operand is the feature map exactly as it lies in HBM
// ([pixel][channel]: "MN-major" for the tensor core) and whose B operand is the softmax-weight matrix [part][pixel]
// (K-major) that the CTA builds once in shared memory.  Split softmax as before: CTA (b, chunk) handles HW / POOL_CHUNKS
// pixels and emits the un-normalised sums together with (m_j, s_j); parthead_kernel merges the chunks.  Unlike the
// mma.sync kernel (elementwise.cu, kept for shapes this one does not take and as ACR_B200_POOL_TC=0) the whole chunk's
// weights are resident next to the feature stages, so there is one maximum per chunk and no rescaling of accumulators.
//
//   warps 0..7  build the weights: thread = pixel, five 16-byte loads cover logit channels 0..39 of its record, the 32
//               part logits go (raw, 16-bit) straight to their place in the K-major SWIZZLE_128B operand; then thread =
//               (part row, eighth of the row): row maximum, exp(l - m) rounded to the storage type in place, s_j = sum
//               of the ROUNDED weights (so the normalised weights still sum to one); generic -> async proxy fence,
//               mbarrier arrive.  Afterwards the same warps are the epilogue: warp & 3 = TMEM lane quadrant, warp >> 2 =
//               channel half; tcgen05.ld 32 columns -> acc[c][0..31] as 128 contiguous bytes per lane.
//   warp 8      TMA producer: per 32-pixel stage one box {64 channels, 32 pixels, 4 channel blocks} (128B swizzle) = the canonical
//               MN-major operand layout: 8 pixel rows x 128 B per atom, atoms of a channel block 1024 B apart (SBO), channel
//               blocks one box (4096 B) apart (LBO).  It starts streaming while the weights are still being built.
//   warp 9      MMA issuer: per stage 2 k-steps x 2 channel halves of tcgen05.mma M=128 N=32 K=16 (A transposed bit set),
//               fp32 accumulators in 64 TMEM columns.
#include <cuda.h>

#include <cstdlib>

#include "ops.cuh"

namespace acr {
namespace {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mb_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {   // bounded: a protocol bug must fail the launch, not hang the box
  for (uint32_t spin = 0; !mb_try(bar, parity); ++spin)
    if (spin > (1u << 26)) { printf("pool_tc: mbarrier timeout (block %d,%d thread %d bar %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar); __trap(); }
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// shared-memory matrix descriptor, SWIZZLE_128B (layout 2), version 1: start >> 4 | LBO >> 4 | SBO >> 4
__device__ __forceinline__ uint64_t desc128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

constexpr int STAGE_PX = 32;                        // pixels per operand stage
constexpr int STAGES = 3;
constexpr int BOX_BYTES = STAGE_PX * 128;           // one {64 channels, 32 pixels} box
constexpr int STAGE_BYTES = 4 * BOX_BYTES;          // 256 channels: 16 KB
constexpr int MAX_PER = 1024;                       // pixels per CTA (HW / POOL_CHUNKS) the resident weights allow
constexpr int W_BYTES = MAX_PER / 64 * 4096;        // [16 k-blocks of 64 pixels][32 part rows][128 B] = 64 KB
constexpr int BUILD_WARPS = 8;
constexpr int THREADS = 32 * (BUILD_WARPS + 2);
constexpr int SMEM = W_BYTES + STAGES * STAGE_BYTES + 256;   // 112.25 KB, no alignment slack: two CTAs per SM fit exactly (2 x (112.25 + 1) <= 228)

struct PoolTcParams {
  CUtensorMap tmF;        // features as {64 channels, B*H*W pixels, 4 channel blocks}, box {64, 32, 4}, 128B swizzle
  const void* logits;     // (B, 2H, 2W, lstride) 16-bit, channel 0 = background
  float* part;
  int HW, W, LW, lstride, per, nkb;
  size_t limg;            // logits elements per image
  uint32_t idesc;
};

template <typename T>
__global__ void __launch_bounds__(THREADS, 2) pool_tc_kernel(const __grid_constant__ PoolTcParams P) {
  extern __shared__ __align__(1024) uint8_t raw_smem[];       // no static shared memory in this kernel: the window starts aligned
  const uint32_t base = s32(raw_smem);
  uint8_t* gen = raw_smem;
  if (base & 1023u) { if (threadIdx.x == 0) printf("pool_tc: dynamic shared memory not 1024-byte aligned (%u)\n", base); __trap(); }
  const uint32_t w_base = base;                               // weights: K-major SWIZZLE_128B, 4 KB per 64-pixel k-block
  const uint32_t f_base = base + W_BYTES;                     // feature stages
  const uint32_t bar_base = f_base + STAGES * STAGE_BYTES;
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t wready = bar_base + 8u * (2 * STAGES), dfull = wready + 8u, tptr = wready + 16u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x, chunk = blockIdx.y, p0 = chunk * P.per;

  if (warp < BUILD_WARPS) {
    // the logit records this thread will transpose: pull all of them towards L2 / L1 now (no registers held), the 16-byte loads
    // below then wait for one memory latency instead of one per batch (ncu: 37 % of the builders' samples sat on those loads)
    const T* lg = static_cast<const T*>(P.logits) + (size_t)b * P.limg;
    for (int pp = threadIdx.x; pp < P.per; pp += 32 * BUILD_WARPS) {
      const int Pg = p0 + pp, y = Pg / P.W, x = Pg - y * P.W;
      const T* rec = lg + ((size_t)(2 * y) * P.LW + 2 * x) * P.lstride;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rec));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + 32));      // channels 32..: the record may straddle a 128-byte line
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mb_init(full(s), 1); mb_init(empty(s), 1); }
    mb_init(wready, 32 * BUILD_WARPS);
    mb_init(dfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == BUILD_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tptr - base));
  float* o = P.part + ((size_t)b * POOL_CHUNKS + chunk) * POOL_PART_FLOATS;
  const int nstages = P.per / STAGE_PX;

  if (warp < BUILD_WARPS) {
    // ========================================================================= softmax weights, then epilogue
    const int t = threadIdx.x;
    const T* lg = static_cast<const T*>(P.logits) + (size_t)b * P.limg;
    // ---- raw logits -> their place in the operand: row = part j, 64 pixels per 128-byte row chunk, 16-byte chunk
    //      ((p & 63) >> 3) ^ (j & 7) (the swizzle follows the address; w_base is 1024-aligned)
    for (int p = t; p < P.per; p += 2 * 32 * BUILD_WARPS) {
      uint32_t w[2][20];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pp = p + u * 32 * BUILD_WARPS;
        if (pp < P.per) {
          const int Pg = p0 + pp, y = Pg / P.W, x = Pg - y * P.W;
          // lanes are consecutive pixels = records 2 * lstride apart: every load instruction touches 32 different lines, so
          // as few (wide) instructions as possible -- five 16-byte loads cover channels 0..39 of the record
          const uint4* rec = reinterpret_cast<const uint4*>(lg + ((size_t)(2 * y) * P.LW + 2 * x) * P.lstride);
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const uint4 r = __ldg(rec + k);
            w[u][4 * k] = r.x; w[u][4 * k + 1] = r.y; w[u][4 * k + 2] = r.z; w[u][4 * k + 3] = r.w;   // channels 2k', 2k' + 1
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pp = p + u * 32 * BUILD_WARPS;
        if (pp < P.per) {
          const uint32_t at = w_base + (uint32_t)(pp >> 6) * 4096u + (uint32_t)(pp & 7) * 2u;
          const uint32_t c = (uint32_t)(pp & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 32; ++j) {                             // part j = channel j + 1
            const uint32_t v = ((j + 1) & 1) ? (w[u][(j + 1) >> 1] >> 16) : (w[u][(j + 1) >> 1] & 0xffffu);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(at + (uint32_t)((j >> 3) * 1024 + (j & 7) * 128) + ((c ^ (uint32_t)(j & 7)) << 4)),
                         "h"((unsigned short)v) : "memory");
          }
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ---- thread = (part row j, eighth e of the row's k-blocks): maximum, then exp in place.  A 16-byte chunk holds eight
    //      consecutive pixels of the row; which eight does not matter here.  Chunk order rotated by e: the eight lanes of a
    //      quarter warp hit eight different bank windows.
    const int j = t >> 3, e = t & 7;
    const uint32_t row = w_base + (uint32_t)((j >> 3) * 1024 + (j & 7) * 128);
    float m = -INFINITY;
    for (int kb = e; kb < P.nkb; kb += 8)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 v;
        const uint32_t a = row + (uint32_t)kb * 4096u + (uint32_t)(((c + e) & 7) << 4);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
        float f[8];
        unpack8<T>(v, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) m = fmaxf(m, f[i]);
      }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    float ssum = 0.f;
    for (int kb = e; kb < P.nkb; kb += 8)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 v;
        const uint32_t a = row + (uint32_t)kb * 4096u + (uint32_t)(((c + e) & 7) << 4);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
        float f[8];
        unpack8<T>(v, f);
        T* q = reinterpret_cast<T*>(&v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          q[i] = from_f32<T>(__expf(f[i] - m));     // ex2.approx path: 2 ulp of fp32, rounded to 16 bits right here
          ssum += to_f32<T>(q[i]);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);
    if (e == 0) { o[256 * 32 + j] = m; o[256 * 32 + 32 + j] = ssum; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core
    mb_arrive(wready);

    // ---- epilogue: lane = channel, 32 columns = parts
    const int q = warp & 3, h = warp >> 2;
    mb_wait(dfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float4* dst = reinterpret_cast<float4*>(o + (size_t)(h * 128 + q * 32 + lane) * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
  } else if (warp == BUILD_WARPS) {
    // ================================================================================== TMA producer
    if (elect_one()) {
      const int gp0 = b * P.HW + p0;
      int s = 0;
      uint32_t ph = 0;
      for (int st = 0; st < nstages; ++st) {
        mb_wait(empty(s), ph ^ 1u);
        mb_expect_tx(full(s), STAGE_BYTES);
        tma_load_3d(f_base + (uint32_t)(s * STAGE_BYTES), &P.tmF, full(s), 0, gp0 + st * STAGE_PX, 0);   // [4 blocks][32 px][128 B]
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // =================================================================================== MMA issuer
    if (elect_one()) {
      mb_wait(wready, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int st = 0; st < nstages; ++st) {
        mb_wait(full(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int ks = 0; ks < STAGE_PX / 16; ++ks) {
          const int px = st * STAGE_PX + ks * 16;                     // first pixel of this k-step inside the chunk
          const uint64_t db = desc128(w_base + (uint32_t)(px >> 6) * 4096u + (uint32_t)(px & 63) * 2u, 16, 1024);
#pragma unroll
          for (int h = 0; h < 2; ++h)     // A: 128 channels = two boxes (LBO), 16 pixels = two 8-row atoms (SBO)
            umma(tmem + (uint32_t)(h * 32), desc128(f_base + (uint32_t)(s * STAGE_BYTES + h * 2 * BOX_BYTES + ks * 2048), BOX_BYTES, 1024), db,
                 P.idesc, (st | ks) ? 1u : 0u);
        }
        commit(empty(s));
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      commit(dfull);
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == BUILD_WARPS + 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

}  // namespace

// ACR_B200_POOL_TC=0 (read at every launch: tests flip it in-process) keeps the mma.sync kernel
bool pool_tc_enabled() {
  const char* e = getenv("ACR_B200_POOL_TC");
  return !(e && atoi(e) == 0);
}

bool pool_tc_takes(const TensorRef& feat, const TensorRef& logits) {
  const int hw = feat.H * feat.W;
  if (feat.C != 256 || hw % POOL_CHUNKS) return false;
  const int per = hw / POOL_CHUNKS;
  return per % 64 == 0 && per <= MAX_PER && feat.pix_stride % 8 == 0 && (uintptr_t)feat.ptr % 16 == 0 && logits.pix_stride % 8 == 0 &&
         logits.pix_stride >= 40 && (uintptr_t)logits.ptr % 16 == 0;
}

int launch_pool_tc(const TensorRef& feat, const TensorRef& logits, float* part, int batch, int act_dtype, cudaStream_t st) {
  ACR_CHECK_ARG(pool_tc_takes(feat, logits) && logits.H == 2 * feat.H && logits.W == 2 * feat.W && logits.C >= 33, "pool_tc: shapes");
  PFN_encodeTiled fn = get_encode();
  if (!fn) { set_error("pool_tc: cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return ACR_B200_ECUDA; }
  PoolTcParams p;
  {
    // channel blocks as the OUTERMOST box dimension although their stride (128 B) is smaller than the pixel stride: one
    // 16 KB box per stage lands as [block][pixel][64 channels] (a TMA box costs ~300-600 clk whatever its size, tools/tma_bench.cu)
    cuuint64_t dims[3] = {64, (cuuint64_t)batch * feat.H * feat.W, 4};
    cuuint64_t str[2] = {(cuuint64_t)feat.pix_stride * 2, 128};
    cuuint32_t box[3] = {64, STAGE_PX, 4}, es[3] = {1, 1, 1};
    const CUtensorMapDataType dt = act_dtype == ACR_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const CUresult r = fn(&p.tmF, dt, 3, feat.ptr, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("pool_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return ACR_B200_ECUDA; }
  }
  p.logits = logits.ptr; p.part = part;
  p.HW = feat.H * feat.W; p.W = feat.W; p.LW = logits.W; p.lstride = logits.pix_stride;
  p.per = p.HW / POOL_CHUNKS; p.nkb = p.per / 64; p.limg = logits.img_stride();
  const uint32_t fmt = act_dtype == ACR_DT_BF16 ? 1u : 0u;
  // D fp32 | A, B format | A is MN-major (bit 15) | N = 32 | M = 128
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (act_dtype == ACR_DT_BF16) {
    static unsigned long long done = 0;
    ACR_CHECK_CUDA(ensure_dynamic_smem(pool_tc_kernel<__nv_bfloat16>, SMEM, &done));
    pool_tc_kernel<__nv_bfloat16><<<dim3(batch, POOL_CHUNKS), THREADS, SMEM, st>>>(p);
  } else if (act_dtype == ACR_DT_F16) {
    static unsigned long long done = 0;
    ACR_CHECK_CUDA(ensure_dynamic_smem(pool_tc_kernel<__half>, SMEM, &done));
    pool_tc_kernel<__half><<<dim3(batch, POOL_CHUNKS), THREADS, SMEM, st>>>(p);
  } else {
    set_error("pool_tc: activation dtype %d", act_dtype);
    return ACR_B200_EINVAL;
  }
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

}  // namespace acr
