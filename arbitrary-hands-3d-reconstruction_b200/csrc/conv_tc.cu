// Implicit-GEMM NHWC convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Replaces every nn.Conv2d + BatchNorm2d (+ReLU, +residual add) of the reference network
// (/root/reference/acr/model.py: BasicBlock :470-499, Bottleneck :501-539, transition / fuse convs
// :620-663, :703-736, SegmNet :374-463, head stacks :288-313, contact conv :227-235), which the
// reference dispatches to cuDNN + separate ATen elementwise kernels.
//
// GEMM view:  D[M = 128 output pixels][N = cout_pad] += A[M][K] * B[N][K],  K = taps * cin_pad.
//   * M tile  = a 16(y) x 8(x) spatial patch of one image; one CTA computes ALL output channels of
//     its patch (N <= 256 fits one UMMA instruction and <= 256 TMEM columns), so activations are
//     never re-read across N tiles.
//   * A operand: for filter tap (ky,kx) and channel chunk c0 the [128 x CK] slice is ONE 4-D TMA box
//     {CK, 8, 16, 1} of the NHWC input at spatial offset (ky-1, kx-1); TMA's out-of-bounds zero
//     fill IS the conv padding.  Stride-2 convs read four parity views (even/odd rows x cols) of the
//     input, each with its own tensor map, so every tap is again a dense box.
//   * B operand: packed weights [cout_pad][taps*cin_pad] (BN folded), one 2-D TMA box {CK, cout_pad}.
//   * Both land in shared memory in the canonical K-major swizzled layout (128B / 64B / 32B swizzle
//     for CK = 64 / 32 / 16) that UMMA shared-memory descriptors address directly.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one thread issues tcgen05.mma, fp32 accumulators
//     in TMEM), warps 2..5 = epilogue (tcgen05.ld -> +bias (+residual) (ReLU) -> 16-bit / fp32 NHWC).
//   * a NUM_STAGES-deep mbarrier ring decouples TMA from the tensor pipe; several CTAs are resident
//     per SM (TMEM columns and smem permitting) so one CTA's epilogue overlaps another's main loop.
#include <cuda.h>

#include "ops.cuh"

namespace acr {

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must fail the launch, not hang the GPU box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) {
      printf("conv_tc: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, bar, parity);
      __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major swizzled operand (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 |
//   [46,48) version=1 | [61,64) layout (2=128B, 4=64B, 6=32B swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// ------------------------------------------------------------------------------------ kernel
constexpr int TILE_Y = 16, TILE_X = 8, TILE_M = 128;
constexpr int TC_THREADS = 192;

struct ConvTcParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  const float* bias;
  const void* res;
  void* out;
  int taps, ksz, stride, cchunks, cin_pad, npad, relu, has_res, out_f32, num_stages, tmem_cols;
  int tiles_x, Ho, Wo, out_stride, res_stride;
  uint32_t idesc;
};

template <int CK>
struct SwizzleCfg {
  static constexpr uint32_t kRowBytes = CK * 2;
  static constexpr uint32_t kSBO = 8 * kRowBytes;  // 8-row core-matrix group
  static constexpr uint32_t kLayout = CK == 64 ? 2u : (CK == 32 ? 4u : 6u);
  static constexpr uint32_t kABytes = TILE_M * kRowBytes;
};

template <int CK, typename T>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ ConvTcParams P) {
  using Cfg = SwizzleCfg<CK>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle atoms need 1024-byte alignment
  const uint32_t b_bytes = (uint32_t)P.npad * Cfg::kRowBytes;
  const uint32_t stage_bytes = (Cfg::kABytes + b_bytes + 1023u) & ~1023u;
  const int S = P.num_stages;
  const uint32_t bar_base = base + (uint32_t)S * stage_bytes;  // full[S], empty[S], tmem_full, tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * S + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ty = blockIdx.x / P.tiles_x, tx = blockIdx.x % P.tiles_x, n = blockIdx.y;
  const int y0 = ty * TILE_Y, x0 = tx * TILE_X;
  const int kiters = P.taps * P.cchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&P.tmB);
    tma_prefetch_desc(&P.tmA[0]);
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < kiters; ++it) {
        const int tap = it / P.cchunks, cc = it % P.cchunks;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), Cfg::kABytes + b_bytes);
        const uint32_t a_dst = base + (uint32_t)s * stage_bytes;
        const uint32_t b_dst = a_dst + Cfg::kABytes;
        int view = 0, dy = 0, dx = 0;
        if (P.ksz == 3) {
          const int ky = tap / 3, kx = tap % 3;
          if (P.stride == 1) { dy = ky - 1; dx = kx - 1; }
          else {  // input row 2*oy + ky - 1 = 2*(oy + dy) + py
            const int py = (ky == 1) ? 0 : 1, px = (kx == 1) ? 0 : 1;
            dy = (ky == 0) ? -1 : 0; dx = (kx == 0) ? -1 : 0;
            view = py * 2 + px;
          }
        }
        tma_load_4d(a_dst, &P.tmA[view], full_bar(s), cc * CK, x0 + dx, y0 + dy, n);
        tma_load_2d(b_dst, &P.tmB, full_bar(s), tap * P.cin_pad + cc * CK, 0);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_src = base + (uint32_t)s * stage_bytes;
        const uint32_t b_src = a_src + Cfg::kABytes;
#pragma unroll
        for (int ks = 0; ks < CK / 16; ++ks) {
          const uint64_t da = make_smem_desc(a_src + ks * 32, Cfg::kSBO, Cfg::kLayout);
          const uint64_t db = make_smem_desc(b_src + ks * 32, Cfg::kSBO, Cfg::kLayout);
          umma_f16(tmem_base, da, db, P.idesc, (it > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ========================================================================= epilogue
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;
    const int oy = y0 + (r >> 3), ox = x0 + (r & 7);
    const size_t pix = ((size_t)n * P.Ho + oy) * P.Wo + ox;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const T* resp = P.has_res ? reinterpret_cast<const T*>(P.res) + pix * P.res_stride : nullptr;
    for (int c0 = 0; c0 < P.npad; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + __ldg(P.bias + c0 + i);
      if (resp) {
        float rr[16];
        unpack8<T>(*reinterpret_cast<const uint4*>(resp + c0), rr);
        unpack8<T>(*reinterpret_cast<const uint4*>(resp + c0 + 8), rr + 8);
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] += rr[i];
      }
      if (P.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      if (P.out_f32) {
        float* o = reinterpret_cast<float*>(P.out) + pix * P.out_stride + c0;
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
      } else {
        T* o = reinterpret_cast<T*>(P.out) + pix * P.out_stride + c0;
        reinterpret_cast<uint4*>(o)[0] = pack8<T>(f);
        reinterpret_cast<uint4*>(o)[1] = pack8<T>(f + 8);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

// --------------------------------------------------------------------------------- host side
struct ConvTcPlan {
  ConvTcParams p;
  int ck, act_dtype, batch, tiles;
  size_t smem;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int encode(CUtensorMap* m, int act_dtype, int rank, const void* ptr, const cuuint64_t* dims,
                  const cuuint64_t* strides, const cuuint32_t* box, int ck) {
  PFN_encodeTiled fn = get_encode();
  if (!fn) { set_error("conv_tc: cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return ACR_B200_ECUDA; }
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = ck == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (ck == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const CUtensorMapDataType dt = act_dtype == ACR_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return ACR_B200_ECUDA; }
  return ACR_B200_OK;
}

int conv_tc_prepare(const ConvArgs& a, int act_dtype, ConvTcPlan** out) {
  ACR_CHECK_ARG(a.out.H % TILE_Y == 0 && a.out.W % TILE_X == 0, "conv_tc: output %dx%d is not a multiple of the 16x8 tile", a.out.H, a.out.W);
  ACR_CHECK_ARG(a.in.pix_stride % 8 == 0 && a.cin_pad % 16 == 0 && a.cout_pad % 16 == 0 && a.cout_pad <= 256,
                "conv_tc: channel alignment");
  ACR_CHECK_ARG(a.in.dtype == act_dtype, "conv_tc: input dtype mismatch");
  ACR_CHECK_ARG(!(a.k == 1 && a.stride != 1), "conv_tc: 1x1 stride-2 unsupported");
  const int ck = (a.cin_pad % 64 == 0) ? 64 : ((a.cin_pad % 32 == 0) ? 32 : 16);
  ConvTcPlan* pl = new ConvTcPlan();
  ConvTcParams& p = pl->p;
  pl->ck = ck; pl->act_dtype = act_dtype; pl->batch = a.batch;
  const cuuint64_t esz = 2;
  const cuuint64_t dim0 = (cuuint64_t)(a.cin_pad < a.in.pix_stride ? a.cin_pad : a.in.pix_stride);
  int rc = ACR_B200_OK;
  if (a.stride == 1) {
    cuuint64_t dims[4] = {dim0, (cuuint64_t)a.in.W, (cuuint64_t)a.in.H, (cuuint64_t)a.batch};
    cuuint64_t str[3] = {(cuuint64_t)a.in.pix_stride * esz, (cuuint64_t)a.in.W * a.in.pix_stride * esz,
                         (cuuint64_t)a.in.H * a.in.W * a.in.pix_stride * esz};
    cuuint32_t box[4] = {(cuuint32_t)ck, TILE_X, TILE_Y, 1};
    rc = encode(&p.tmA[0], act_dtype, 4, a.in.ptr, dims, str, box, ck);
    for (int v = 1; v < 4 && !rc; ++v) p.tmA[v] = p.tmA[0];
  } else {
    for (int v = 0; v < 4 && !rc; ++v) {
      const int py = v >> 1, px = v & 1;
      const char* ptr = static_cast<const char*>(a.in.ptr) + ((size_t)py * a.in.W + px) * a.in.pix_stride * esz;
      cuuint64_t dims[4] = {dim0, (cuuint64_t)a.in.W / 2, (cuuint64_t)a.in.H / 2, (cuuint64_t)a.batch};
      cuuint64_t str[3] = {(cuuint64_t)2 * a.in.pix_stride * esz, (cuuint64_t)2 * a.in.W * a.in.pix_stride * esz,
                           (cuuint64_t)a.in.H * a.in.W * a.in.pix_stride * esz};
      cuuint32_t box[4] = {(cuuint32_t)ck, TILE_X, TILE_Y, 1};
      rc = encode(&p.tmA[v], act_dtype, 4, ptr, dims, str, box, ck);
    }
  }
  if (!rc) {
    const int taps = a.k * a.k;
    cuuint64_t dims[2] = {(cuuint64_t)taps * a.cin_pad, (cuuint64_t)a.cout_pad};
    cuuint64_t str[1] = {(cuuint64_t)taps * a.cin_pad * esz};
    cuuint32_t box[2] = {(cuuint32_t)ck, (cuuint32_t)a.cout_pad};
    rc = encode(&p.tmB, act_dtype, 2, a.w, dims, str, box, ck);
  }
  if (rc) { delete pl; return rc; }
  p.bias = a.bias; p.res = a.has_res ? a.res.ptr : nullptr; p.out = a.out.ptr;
  p.taps = a.k * a.k; p.ksz = a.k; p.stride = a.stride; p.cchunks = a.cin_pad / ck; p.cin_pad = a.cin_pad;
  p.npad = a.cout_pad; p.relu = a.relu; p.has_res = a.has_res; p.out_f32 = a.out.dtype == ACR_DT_F32;
  p.tiles_x = a.out.W / TILE_X; p.Ho = a.out.H; p.Wo = a.out.W; p.out_stride = a.out.pix_stride;
  p.res_stride = a.has_res ? a.res.pix_stride : 0;
  p.tmem_cols = 32;
  while (p.tmem_cols < a.cout_pad) p.tmem_cols *= 2;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A/B = bf16|f16, K-major both, N, M=128
  const uint32_t fmt = act_dtype == ACR_DT_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a.cout_pad >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  const size_t stage = (((size_t)TILE_M * ck * 2 + (size_t)a.cout_pad * ck * 2) + 1023) & ~(size_t)1023;
  const int kiters = p.taps * p.cchunks;
  int S = kiters < 4 ? kiters : 4;
  while (S > 1 && S * stage + 2048 > 200 * 1024) --S;
  p.num_stages = S;
  pl->smem = S * stage + 1024 /*alignment slack*/ + 8 * (2 * S + 2) + 64;
  pl->tiles = p.tiles_x * (a.out.H / TILE_Y);
  *out = pl;
  return ACR_B200_OK;
}

template <int CK, typename T>
static int launch_inst(const ConvTcPlan* pl, cudaStream_t st) {
  static size_t configured = 0;
  if (pl->smem > configured) {
    ACR_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CK, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    configured = 220 * 1024;
  }
  conv_tc_kernel<CK, T><<<dim3(pl->tiles, pl->batch), TC_THREADS, pl->smem, st>>>(pl->p);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

int conv_tc_launch(const ConvTcPlan* pl, cudaStream_t st) {
  const bool bf = pl->act_dtype == ACR_DT_BF16;
  switch (pl->ck) {
    case 64: return bf ? launch_inst<64, __nv_bfloat16>(pl, st) : launch_inst<64, __half>(pl, st);
    case 32: return bf ? launch_inst<32, __nv_bfloat16>(pl, st) : launch_inst<32, __half>(pl, st);
    default: return bf ? launch_inst<16, __nv_bfloat16>(pl, st) : launch_inst<16, __half>(pl, st);
  }
}

void conv_tc_free(ConvTcPlan* p) { delete p; }

}  // namespace acr
