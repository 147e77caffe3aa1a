// Stem conv on the tensor cores WITHOUT the im2col round trip (sm_100a).
//
// HigherResolutionNet.forward /root/reference/acr/model.py:832-835: x/255*2-1, conv1 3x3 stride 2 (3 -> 64) + bn1 + ReLU.
// Round 1 ran it as im2col_stem_kernel (uint8 frame -> 32-channel 16-bit tensor of the 27 normalised taps, 1.07 GB at
// batch 256) + a 1x1 tcgen05 conv that reads it back.  Here the A operand of that GEMM is built IN SHARED MEMORY by
// producer warps straight from the uint8 frame, in the K-major SWIZZLE_64B layout a TMA box {32, 16, 16} would have
// produced (row = output pixel, 32 channels = 27 taps + 5 zeros = 64 bytes; 16-byte chunk c of row r lives at chunk
// c ^ ((r >> 1) & 3): the swizzle is a function of the absolute shared-memory address, tools/umma_shift_probe.cu), so
// the intermediate tensor never exists: 0.2 GB in + 2.15 GB out instead of 0.2 + 1.07 + 1.07 + 2.15 GB.
//
// Persistent CTAs, one per SM, tile = 16x16 output pixels (two M = 128 UMMA tiles: left / right 8 columns):
//   warps 0..7   epilogue: tcgen05.ld -> +bias -> ReLU -> 16-bit NHWC, 256-bit stores        (warp & 3 = TMEM lane quadrant)
//   warp  8      MMA issuer: one elected thread, 2 k-steps x 2 halves of M=128 N=64 K=16 per tile, fp32 accumulators in TMEM
//   warps 9..16  producers: thread = output pixel; stage the 33x33 uint8 patch of the tile in shared memory (coalesced),
//                look the 27 taps up in a 256-entry table of (float)v / 255.f * 2.f - 1.f (bit-identical to the reference's
//                normalisation), round to the storage type, write the swizzled 64-byte row; generic -> async proxy fence,
//                producer-only named barrier, one mbarrier arrive per stage.
#include <cuda.h>

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
    if (spin > (1u << 26)) { printf("stem_tc: mbarrier timeout (block %d thread %d bar %u)\n", blockIdx.x, threadIdx.x, bar); __trap(); }
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
// K-major SWIZZLE_64B descriptor: start >> 4 | LBO (unused) | SBO | version 1 | layout 4
__device__ __forceinline__ uint64_t desc64(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

constexpr int NS = 4;                       // operand stages
constexpr int EPI = 8, PROD_WARPS = 8;
constexpr int THREADS = 32 * (EPI + 1 + PROD_WARPS);
constexpr int PATCH_ROW = 112;              // bytes per staged patch row (33 pixels x 3 = 99, padded)
constexpr int A_BYTES = 256 * 64;           // one stage of the A operand: 256 pixel rows x 64 B
constexpr int PATCH_BYTES = 33 * PATCH_ROW;
constexpr int SMEM = 1024 + NS * A_BYTES + 4096 /*weights*/ + NS * PATCH_BYTES + 1024 /*lut*/ + 256 /*bias*/ + 256 /*barriers*/;

struct StemTcParams {
  const uint8_t* img;     // (B, H, W, 3) uint8
  void* out;              // (B, H/2, W/2, out_stride) 16-bit
  const void* w;          // packed [64][32] 16-bit, K-major (channel (ky*3+kx)*3+ci, 27..31 zero), BN folded
  const float* bias;      // [64]
  int H, W, out_stride, tiles_x, tiles_per_img, total_tiles;
  uint32_t idesc;
};

template <typename T>
__global__ void __launch_bounds__(THREADS, 1) stem_tc_kernel(const __grid_constant__ StemTcParams P) {
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = s32(raw_smem);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = raw_smem + (base - raw);                 // generic pointer to the aligned base
  const uint32_t a_base = base;
  const uint32_t w_base = base + NS * A_BYTES;
  uint8_t* patch = gen + NS * A_BYTES + 4096;
  float* lut = reinterpret_cast<float*>(gen + NS * A_BYTES + 4096 + NS * PATCH_BYTES);
  float* s_bias = lut + 256;
  const uint32_t bar_base = s32(s_bias + 64);
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (NS + s); };
  auto tfull = [&](int b) { return bar_base + 8u * (2 * NS + b); };
  auto tempty = [&](int b) { return bar_base + 8u * (2 * NS + 2 + b); };
  const uint32_t tptr = bar_base + 8u * (2 * NS + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mb_init(fullA(s), 1); mb_init(emptyA(s), 1); }
    for (int b = 0; b < 2; ++b) { mb_init(tfull(b), 1); mb_init(tempty(b), EPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 256) lut[threadIdx.x] = (float)threadIdx.x / 255.f * 2.0f - 1.0f;   // the reference's fp32 normalisation
  if (threadIdx.x < 64) s_bias[threadIdx.x] = P.bias[threadIdx.x];
  // weights [64 rows][64 B] into the SWIZZLE_64B layout: 256 16-byte chunks
  if (threadIdx.x < 256) {
    const int r = threadIdx.x >> 2, c = threadIdx.x & 3;
    const uint4 v = reinterpret_cast<const uint4*>(P.w)[threadIdx.x];
    *reinterpret_cast<uint4*>(gen + NS * A_BYTES + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = v;
  }
  if (warp == EPI) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the weights were written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tptr - base));

  if (warp < EPI) {
    // ===================================================================================== epilogue
    const int q = warp & 3, h = warp >> 2, r = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)it >> 1;
      const int n = tile / P.tiles_per_img, rem = tile % P.tiles_per_img;
      const int oy = (rem / P.tiles_x) * 16 + (r >> 3), ox = (rem % P.tiles_x) * 16 + h * 8 + (r & 7);
      const size_t pix = ((size_t)n * (P.H / 2) + oy) * (P.W / 2) + ox;
      mb_wait(tfull(buf), use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + h) * 64);
      uint32_t v[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c) ld16(t_row + (uint32_t)(c * 16), v[c]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mb_arrive(tempty(buf));
      T* o = reinterpret_cast<T*>(P.out) + pix * P.out_stride;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fmaxf(__uint_as_float(v[c][i]) + s_bias[c * 16 + i], 0.f);
        stg256(o + c * 16, pack8<T>(f), pack8<T>(f + 8));
      }
    }
  } else if (warp == EPI) {
    // =================================================================================== MMA issuer
    if (elect_one()) {
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)it >> 1;
        mb_wait(tempty(buf), (use & 1u) ^ 1u);
        mb_wait(fullA(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = a_base + (uint32_t)s * A_BYTES;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)     // rows of a half tile: 16 groups of 8 pixels, one image row (16 px x 64 B) apart
            umma(tmem + (uint32_t)((buf * 2 + hh) * 64), desc64(a0 + hh * 512 + ks * 32, 1024), desc64(w_base + ks * 32, 512), P.idesc, ks ? 1u : 0u);
        commit(emptyA(s));
        commit(tfull(buf));
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ===================================================================================== producers
    const int t = threadIdx.x - 32 * (EPI + 1);     // 0..255 = output pixel of the tile
    const int py = t >> 4, px = t & 15;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const int n = tile / P.tiles_per_img, rem = tile % P.tiles_per_img;
      const int y0 = (rem / P.tiles_x) * 16, x0 = (rem % P.tiles_x) * 16;
      const int iy0 = 2 * y0 - 1, ix0 = 2 * x0 - 1;                       // input pixel of patch (0, 0)
      mb_wait(emptyA(s), ph ^ 1u);                                         // the MMAs that read this stage are done
      uint8_t* pp = patch + s * PATCH_BYTES;
      // ---- stage the 33 x 33 x 3 uint8 patch (rows of 99 contiguous bytes of the frame)
      const uint8_t* img = P.img + (size_t)n * P.H * P.W * 3;
      for (int i = t; i < 33 * 25; i += 256) {                             // 25 4-byte words per row (100 >= 99 bytes)
        const int r = i / 25, wq = i - r * 25;
        const int iy = iy0 + r;
        uint32_t word = 0;
        if (iy >= 0 && iy < P.H) {
          const long long rowoff = ((long long)iy * P.W) * 3;
          const int b0 = ix0 * 3 + wq * 4;                                 // byte offset inside the image row (may be < 0)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int bb = b0 + k;
            const uint32_t v = (bb >= 0 && bb < P.W * 3) ? (uint32_t)img[rowoff + bb] : 0u;
            word |= v << (8 * k);
          }
        }
        *reinterpret_cast<uint32_t*>(pp + r * PATCH_ROW + wq * 4) = word;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");                        // producers only
      // ---- this thread's pixel: 27 taps -> 32 channels -> one swizzled 64-byte row
      float vch[32];
#pragma unroll
      for (int i = 27; i < 32; ++i) vch[i] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = iy0 + 2 * py + ky;
        const bool yok = iy >= 0 && iy < P.H;
        const uint8_t* prow = pp + (2 * py + ky) * PATCH_ROW + (2 * px) * 3;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ix0 + 2 * px + kx;
          const bool ok = yok && ix >= 0 && ix < P.W;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) vch[(ky * 3 + kx) * 3 + ci] = ok ? lut[prow[kx * 3 + ci]] : 0.f;
        }
      }
      const int r = py * 16 + px;
      uint8_t* arow = gen + s * A_BYTES + r * 64;
      const int sw = (r >> 1) & 3;
#pragma unroll
      for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(arow + ((c ^ sw) << 4)) = pack8<T>(vch + c * 8);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic-proxy writes -> visible to the tensor core
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (t == 0) mb_arrive(fullA(s));
      if (++s == NS) { s = 0; ph ^= 1u; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == EPI) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

}  // namespace

int launch_stem_tc(const TensorRef& img, const TensorRef& out, const void* w, const float* bias, int batch, int act_dtype,
                   cudaStream_t st) {
  ACR_CHECK_ARG(out.C == 64 && out.H * 2 == img.H && out.W * 2 == img.W && img.dtype == ACR_DT_U8 && out.H % 16 == 0 &&
                    out.W % 16 == 0 && out.pix_stride % 16 == 0 && (uintptr_t)out.ptr % 32 == 0 && (uintptr_t)w % 16 == 0 &&
                    out.dtype == act_dtype, "stem_tc: shape / alignment");
  StemTcParams p;
  p.img = static_cast<const uint8_t*>(img.ptr); p.out = out.ptr; p.w = w; p.bias = bias;
  p.H = img.H; p.W = img.W; p.out_stride = out.pix_stride;
  p.tiles_x = out.W / 16; p.tiles_per_img = p.tiles_x * (out.H / 16); p.total_tiles = p.tiles_per_img * batch;
  const uint32_t fmt = act_dtype == ACR_DT_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (act_dtype == ACR_DT_BF16) {
    static unsigned long long done = 0;
    ACR_CHECK_CUDA(ensure_dynamic_smem(stem_tc_kernel<__nv_bfloat16>, SMEM, &done));
    stem_tc_kernel<__nv_bfloat16><<<grid, THREADS, SMEM, st>>>(p);
  } else if (act_dtype == ACR_DT_F16) {
    static unsigned long long done = 0;
    ACR_CHECK_CUDA(ensure_dynamic_smem(stem_tc_kernel<__half>, SMEM, &done));
    stem_tc_kernel<__half><<<grid, THREADS, SMEM, st>>>(p);
  } else {
    set_error("stem_tc: activation dtype %d", act_dtype);
    return ACR_B200_EINVAL;
  }
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

}  // namespace acr
