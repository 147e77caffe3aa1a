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
//   warps 0..7   epilogue (warp & 3 = TMEM lane quadrant, warp >> 2 = half): tcgen05.ld -> 16-bit pairs -> ReLU ->
//                128B-swizzled 4 KB slab in shared memory (32 pixels x 64 channels) -> TMA store, two tiles in flight per warp
//   warp  8      MMA issuer: one elected thread, 2 k-steps x 2 halves of M=128 N=64 K=16 per tile, fp32 accumulators in TMEM
//   warps 9..24  producers, two groups of 8 warps taking alternate tiles (a group converts one tile at a time, so one group
//                alone leaves the load latency exposed): thread = output pixel; its 27 uint8 taps straight from the frame
//                (L1 shares them between neighbours; the next tile's patch is prefetched), (float)v / 255.f * 2.f - 1.f
//                in three FMA-pipe operations (bit-identical to the reference's normalisation), round to the storage
//                type, write the swizzled 64-byte row; generic -> async proxy fence, one mbarrier arrive per warp.
// The BN bias rides in two spare K channels (hi + lo 16-bit parts against constant-one taps); ReLU on the packed 16-bit pair.
// Measured at batch 256 (B200): 0.48 ms against 0.50 ms (im2col) + 0.58 ms (1x1 conv); profiles/r2_bench_ab_stem_fused.json,
// profiles/r2_stem_ncu_summary.md (4.85 TB/s of DRAM traffic = 0.74 of the copy peak).
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
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) { asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// (float)b / 255.f * 2.f - 1.f for a byte b, bit for bit (the reference's normalisation, acr/model.py:832), without a division
// or a table: b as a float through the 2^23 trick, the correctly rounded quotient as fma(x, rh, x * rl) with rh + rl = 1/255
// to 48 bits (exact for all 256 inputs: tests/test_cpu_round2.py), then 2q - 1 in one rounding (2q is exact).
__device__ __forceinline__ float normalised(uint32_t b) {
  const float x = __uint_as_float(0x4B000000u | b) - 8388608.f;
  const float q = __fmaf_rn(x, __uint_as_float(0x3B808081u), __fmul_rn(x, __uint_as_float(0xAF7EFEFFu)));
  return __fmaf_rn(q, 2.f, -1.f);
}
// two fp32 -> packed 16-bit pair (round to nearest even), ReLU on the packed pair: rounding is monotonic and keeps zero, so
// max(round(x), 0) == round(max(x, 0)); half the instructions of fmaxf + convert
template <typename T> struct Pair16;
template <> struct Pair16<__nv_bfloat16> { using t = __nv_bfloat162; static __device__ __forceinline__ t cvt(float a, float b) { return __floats2bfloat162_rn(a, b); } };
template <> struct Pair16<__half> { using t = __half2; static __device__ __forceinline__ t cvt(float a, float b) { return __floats2half2_rn(a, b); } };
template <typename T>
__device__ __forceinline__ uint4 relu_pack8(const uint32_t* v) {
  uint4 u;
  typename Pair16<T>::t* p = reinterpret_cast<typename Pair16<T>::t*>(&u);
  const typename Pair16<T>::t zero = Pair16<T>::cvt(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __hmax2(Pair16<T>::cvt(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), zero);
  return u;
}
// K-major SWIZZLE_64B descriptor: start >> 4 | LBO (unused) | SBO | version 1 | layout 4
__device__ __forceinline__ uint64_t desc64(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

constexpr int NS = 4;                       // operand stages
constexpr int EPI = 8, PROD_WARPS = 16;      // two producer groups of 8 warps take alternate tiles
constexpr int THREADS = 32 * (EPI + 1 + PROD_WARPS);
constexpr int A_BYTES = 256 * 64;           // one stage of the A operand: 256 pixel rows x 64 B
constexpr int SLAB = 4096;                  // one epilogue warp's 32 pixels x 64 channels, the box of a TMA store
constexpr int OUT_BYTES = EPI * 2 * SLAB;   // two slabs (tiles in flight) per epilogue warp
constexpr int SMEM = 1024 + NS * A_BYTES + 4096 /*weights*/ + OUT_BYTES + 256 /*barriers*/;

struct StemTcParams {
  CUtensorMap tmOut;      // output as {64, W/2, H/2, B}, box {64, 8, 4, 1}, 128B swizzle
  const uint8_t* img;     // (B, H, W, 3) uint8
  void* out;              // (B, H/2, W/2, out_stride) 16-bit
  const void* w;          // packed [64][32] 16-bit, K-major (channel (ky*3+kx)*3+ci, 27..31 zero), BN folded
  const float* bias;      // [64]
  int H, W, out_stride, total_tiles;
  int tx_log2, tpi_log2;  // tiles per output row / per image are powers of two (512 x 512 frames: 16, 256): shifts, no divisions in the tile loops
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
  const uint32_t out_base = base + NS * A_BYTES + 4096;     // 1024-aligned: the 128B swizzle of the slabs follows the address
  const uint32_t bar_base = base + NS * A_BYTES + 4096 + OUT_BYTES;
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (NS + s); };
  auto tfull = [&](int b) { return bar_base + 8u * (2 * NS + b); };
  auto tempty = [&](int b) { return bar_base + 8u * (2 * NS + 2 + b); };
  const uint32_t tptr = bar_base + 8u * (2 * NS + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mb_init(fullA(s), 8); mb_init(emptyA(s), 1); }   // one arrival per producer warp of a group
    for (int b = 0; b < 2; ++b) { mb_init(tfull(b), 1); mb_init(tempty(b), EPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights [64 rows][64 B] into the SWIZZLE_64B layout: 256 16-byte chunks
  if (threadIdx.x < 256) {
    const int r = threadIdx.x >> 2, c = threadIdx.x & 3;
    uint4 v = reinterpret_cast<const uint4*>(P.w)[threadIdx.x];
    if (c == 3) {   // the bias rides in two of the five spare K channels (the producers write 1.0 there): hi + lo parts, so the
                    // fp32 accumulator receives it to 2^-17 (bf16) / 2^-22 (fp16) relative and the epilogue needs no add
      T* e = reinterpret_cast<T*>(&v);
      const float b = P.bias[r];
      e[3] = from_f32<T>(b);
      e[4] = from_f32<T>(b - to_f32<T>(e[3]));
    }
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
    const int q = warp & 3, h = warp >> 2;               // TMEM lane quadrant = rows 4q .. 4q+3 of the tile; left / right half
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)it >> 1;
      const int n = tile >> P.tpi_log2, rem = tile & ((1 << P.tpi_log2) - 1);
      const int ty = (rem >> P.tx_log2) * 16 + 4 * q, tx = (rem & ((1 << P.tx_log2) - 1)) * 16 + h * 8;   // slab origin
      mb_wait(tfull(buf), use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + h) * 64);
      uint32_t v[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c) ld16(t_row + (uint32_t)(c * 16), v[c]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mb_arrive(tempty(buf));
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store of tile it - 2 has read this slab
      }
      __syncwarp();
      // lane = pixel (row lane >> 3, column lane & 7) of the slab = 128-byte line `lane`: chunk c lives at c ^ (lane & 7)
      const uint32_t slab = out_base + (uint32_t)(warp * 2 + buf) * SLAB, row = slab + (uint32_t)lane * 128u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {      // the accumulator already holds conv + bias: ReLU, round, store
        sts128(row + (((uint32_t)(2 * c) ^ (uint32_t)(lane & 7)) << 4), relu_pack8<T>(v[c]));
        sts128(row + (((uint32_t)(2 * c + 1) ^ (uint32_t)(lane & 7)) << 4), relu_pack8<T>(v[c] + 8));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the TMA unit
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&P.tmOut, slab, 0, tx, ty, n);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the slabs must outlive the stores
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
    const int pt = threadIdx.x - 32 * (EPI + 1);    // 0..511
    const int grp = pt >> 8, t = pt & 255;          // group 0 / 1 builds the even / odd tiles of this CTA; t = output pixel
    const int py = t >> 4, px = t & 15;
    int s = grp;                                    // stage of tile number it = it % NS; this group's tiles: it = grp, grp + 2, ..
    uint32_t ph = 0;
    const int r = py * 16 + px;
    const int sw = (r >> 1) & 3;
    const int tmask = (1 << P.tpi_log2) - 1, xmask = (1 << P.tx_log2) - 1;
    const size_t img_bytes = (size_t)P.H * P.W * 3;
    // Only the top and the left frame edges pad (the patch of tile (y0, x0) starts at input pixel (2*y0 - 1, 2*x0 - 1) and ends
    // inside the frame): filter row 0 of pixel row 0 of the tiles with y0 == 0, filter column 0 of pixel column 0 where x0 == 0.
    for (int tile = blockIdx.x + grp * (int)gridDim.x; tile < P.total_tiles; tile += 2 * (int)gridDim.x) {
      const int n = tile >> P.tpi_log2, rem = tile & tmask;
      const int y0 = (rem >> P.tx_log2) * 16, x0 = (rem & xmask) * 16;
      const bool pad_top = (y0 | py) == 0, pad_left = (x0 | px) == 0;
      const int iy = 2 * (y0 + py) - (pad_top ? 0 : 1), ix = 2 * (x0 + px) - (pad_left ? 0 : 1);   // first tap that exists
      const uint8_t* p0 = P.img + (size_t)n * img_bytes + ((size_t)iy * P.W + ix) * 3;
      {  // pull the NEXT tile's 33 x 99-byte patch towards L1 while this one is converted
        const int nt = tile + 2 * (int)gridDim.x;
        if (nt < P.total_tiles && t < 66) {
          const int nrem = nt & tmask;
          int piy = 2 * (nrem >> P.tx_log2) * 16 - 1 + (t >> 1), pix = 2 * (nrem & xmask) * 16 - 1;
          piy = piy < 0 ? 0 : piy; pix = pix < 0 ? 0 : pix;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(P.img + (size_t)(nt >> P.tpi_log2) * img_bytes + ((size_t)piy * P.W + pix) * 3 + (t & 1) * 96));
        }
      }
      mb_wait(emptyA(s), ph ^ 1u);                                         // the MMAs that read this stage are done
      // ---- this thread's pixel: 3 filter rows x 9 contiguous bytes straight from the frame (neighbouring pixels share them
      // through L1); a padded row / column reads the next one instead (always inside the frame) and is zeroed afterwards
      uint32_t raw9[3][9];
      if (x0 != 0) {        // (uniform) no left padding: the row starts at an odd address -> one byte + four aligned 16-bit loads
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint8_t* prow = p0 + (size_t)(pad_top ? (ky ? ky - 1 : 0) : ky) * P.W * 3;
          raw9[ky][0] = (uint32_t)__ldg(prow);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(prow + 1 + 2 * j));
            raw9[ky][1 + 2 * j] = w & 0xffu;
            raw9[ky][2 + 2 * j] = w >> 8;
          }
        }
      } else {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint8_t* prow = p0 + (size_t)(pad_top ? (ky ? ky - 1 : 0) : ky) * P.W * 3 - (pad_left ? 3 : 0);
          const uint8_t* pcol0 = prow + (pad_left ? 3 : 0);                 // filter column 0 (or its stand-in)
#pragma unroll
          for (int j = 0; j < 9; ++j) raw9[ky][j] = (uint32_t)__ldg((j < 3 ? pcol0 : prow) + j);
        }
      }
      float vch[32];
#pragma unroll
      for (int i = 27; i < 32; ++i) vch[i] = i < 29 ? 1.f : 0.f;           // channels 27 / 28 carry the bias (hi / lo)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const float v = normalised(raw9[ky][j]);
          vch[ky * 9 + j] = (ky == 0 && j < 3) ? ((pad_top || pad_left) ? 0.f : v) : ky == 0 ? (pad_top ? 0.f : v) : j < 3 ? (pad_left ? 0.f : v) : v;
        }
      uint8_t* arow = gen + s * A_BYTES + r * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(arow + ((c ^ sw) << 4)) = pack8<T>(vch + c * 8);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mb_arrive(fullA(s));                                  // 8 arrivals (this group's warps) complete the stage
      s += 2;
      if (s >= NS) { s -= NS; ph ^= 1u; }
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
                    out.dtype == act_dtype && (uintptr_t)img.ptr % 2 == 0, "stem_tc: shape / alignment");
  StemTcParams p;
  {
    const int rc = encode_slab_store_map(&p.tmOut, out, 64, batch, act_dtype);
    if (rc) return rc;
  }
  p.img = static_cast<const uint8_t*>(img.ptr); p.out = out.ptr; p.w = w; p.bias = bias;
  p.H = img.H; p.W = img.W; p.out_stride = out.pix_stride;
  const int tiles_x = out.W / 16, tiles_per_img = tiles_x * (out.H / 16);
  ACR_CHECK_ARG((tiles_x & (tiles_x - 1)) == 0 && (tiles_per_img & (tiles_per_img - 1)) == 0, "stem_tc: tiles per row / image must be powers of two");
  p.tx_log2 = __builtin_ctz(tiles_x); p.tpi_log2 = __builtin_ctz(tiles_per_img); p.total_tiles = tiles_per_img * batch;
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
