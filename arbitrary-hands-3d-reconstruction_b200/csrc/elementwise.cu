// Memory-bound helper kernels of the ACR network (everything that is not an implicit-GEMM conv):
// stem conv on uint8 input, HRNet fuse-sum, bilinear x2, coord channels, attention pooling
// (Hadamard_product), the per-image part head, the folded final 1x1 conv, and a CUDA-core
// reference conv used by the tests to localise tensor-core bugs.
// Reference call sites are cited per kernel (/root/reference/acr/model.py unless noted).
#include "ops.cuh"

namespace acr {

#define ACR_DISPATCH_ACT(dt, ...)                                        \
  do {                                                                   \
    if ((dt) == ACR_DT_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }   \
    else if ((dt) == ACR_DT_F16) { using T = __half; __VA_ARGS__; }      \
    else { set_error("unsupported activation dtype %d", (int)(dt)); return ACR_B200_EINVAL; } \
  } while (0)

// ------------------------------------------------------------------------------------ stem
// HigherResolutionNet.forward :832-835: x/255*2-1, conv1 3x3 s2 (3->64) + bn1 + relu.
// weights: fp32 [27][64] (tap-major, BN folded), bias fp32 [64].  thread = one output pixel x all 64
// channels: the 27 normalised inputs are scalars, the 64 weights of a tap are broadcast float4 reads
// from shared memory, 64 fp32 accumulators live in registers (1728 FMA per 4 B of input).
template <typename T>
__global__ void __launch_bounds__(128) stem_kernel(const uint8_t* __restrict__ img, T* __restrict__ out,
                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                   int H, int W, int out_stride, long long total) {
  __shared__ __align__(16) float s_w[27 * 64];
  __shared__ __align__(16) float s_b[64];
  for (int i = threadIdx.x; i < 27 * 64; i += 128) s_w[i] = w[i];
  if (threadIdx.x < 64) s_b[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * 128 + threadIdx.x;
  if (pix >= total) return;
  const int Ho = H / 2, Wo = W / 2;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho);
  const long long b = pix / ((long long)Wo * Ho);
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; c += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(&s_b[c]);
    acc[c] = b4.x; acc[c + 1] = b4.y; acc[c + 2] = b4.z; acc[c + 3] = b4.w;
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - 1;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - 1;
      if (ix < 0 || ix >= W) continue;
      const uint8_t* px = img + ((b * H + iy) * W + ix) * 3;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float xn = (float)px[ci] / 255.f * 2.0f - 1.0f;
        const float4* wr = reinterpret_cast<const float4*>(&s_w[((ky * 3 + kx) * 3 + ci) * 64]);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float4 w4 = wr[c];
          acc[4 * c + 0] = fmaf(xn, w4.x, acc[4 * c + 0]); acc[4 * c + 1] = fmaf(xn, w4.y, acc[4 * c + 1]);
          acc[4 * c + 2] = fmaf(xn, w4.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(xn, w4.w, acc[4 * c + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = fmaxf(acc[c], 0.f);
  T* o = out + pix * out_stride;
#pragma unroll
  for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(o + c * 8) = pack8<T>(acc + c * 8);
}

int launch_stem(const TensorRef& img, const TensorRef& out, const float* w, const float* bias, int batch,
                int act_dtype, cudaStream_t st) {
  ACR_CHECK_ARG(out.C == 64 && out.H * 2 == img.H && out.W * 2 == img.W && img.dtype == ACR_DT_U8,
                "stem: shape mismatch");
  const long long total = (long long)batch * out.H * out.W;
  ACR_DISPATCH_ACT(act_dtype, stem_kernel<T><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
                                  (const uint8_t*)img.ptr, (T*)out.ptr, w, bias, img.H, img.W, out.pix_stride, total));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------ im2col stem
// The stem conv (acr/model.py:832-835) on the tensor cores: this kernel only gathers the 27 normalised taps
// of every stride-2 output pixel into a 32-channel 16-bit tensor (channel (ky*3+kx)*3+ci, 0 for taps in the
// conv padding and for channels 27..31); the 27->64 contraction + BN + ReLU is then a 1x1 conv_tc launch.
template <typename T>
__global__ void __launch_bounds__(256) im2col_stem_kernel(const uint8_t* __restrict__ img, T* __restrict__ out,
                                                          int H, int W, int out_stride) {
  // grid = (ceil(Wo/256), Ho, B): no integer divisions on the address path
  const int Ho = H / 2, Wo = W / 2;
  const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y;
  if (ox >= Wo) return;
  const size_t b = blockIdx.z;
  float v[32];
#pragma unroll
  for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - 1;
    const bool yok = iy >= 0 && iy < H;
    const uint8_t* row = img + ((b * H + (yok ? iy : 0)) * (size_t)W) * 3;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - 1;
      const bool ok = yok && ix >= 0 && ix < W;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        v[(ky * 3 + kx) * 3 + ci] = ok ? (float)row[ix * 3 + ci] / 255.f * 2.0f - 1.0f : 0.f;
    }
  }
  T* o = out + ((b * Ho + oy) * (size_t)Wo + ox) * out_stride;
  stg256(o, pack8<T>(v), pack8<T>(v + 8));
  stg256(o + 16, pack8<T>(v + 16), pack8<T>(v + 24));
}

int launch_im2col_stem(const TensorRef& img, const TensorRef& out, int batch, int act_dtype, cudaStream_t st) {
  ACR_CHECK_ARG(out.C == 32 && out.H * 2 == img.H && out.W * 2 == img.W && img.dtype == ACR_DT_U8 && out.pix_stride >= 32 &&
                    out.pix_stride % 16 == 0 && (uintptr_t)out.ptr % 32 == 0 && out.H <= 65535 && batch <= 65535, "im2col_stem: shape mismatch");
  const dim3 grid((unsigned)((out.W + 255) / 256), (unsigned)out.H, (unsigned)batch);
  ACR_DISPATCH_ACT(act_dtype, im2col_stem_kernel<T><<<grid, 256, 0, st>>>((const uint8_t*)img.ptr, (T*)out.ptr, img.H, img.W,
                                                                           out.pix_stride));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ---------------------------------------------------------------------------- reference conv
// Same contract as the tcgen05 conv (ConvArgs): NHWC, weights [cout_pad][k*k][cin_pad], fp32
// accumulate, epilogue = +bias (+residual) (ReLU), 16-bit or fp32 NHWC output.
// thread = (output pixel, 8 output channels).  Debug / parity tool only.
template <typename T>
__global__ void __launch_bounds__(128) conv_ref_kernel(ConvArgs a, long long total) {
  const long long gid = (long long)blockIdx.x * 128 + threadIdx.x;
  if (gid >= total) return;
  const int ngrp = a.cout_pad / 8;
  const int cg = (int)(gid % ngrp);
  const long long pix = gid / ngrp;
  const int Wo = a.out.W, Ho = a.out.H;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho);
  const int b = (int)(pix / ((long long)Wo * Ho));
  const T* in = (const T*)a.in.ptr + (size_t)b * a.in.img_stride();
  const T* w = (const T*)a.w;
  const int pad = a.k / 2, taps = a.k * a.k;
  float acc[8];
  const float* bias = a.bias + (a.bias_per_image ? (size_t)b * a.cout_pad : 0) + cg * 8;
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = bias[c];
  const int cin_vec = (a.in.C + 7) / 8;  // channels physically present (multiple of 8)
  for (int ky = 0; ky < a.k; ++ky) {
    const int iy = oy * a.stride + ky - pad;
    if (iy < 0 || iy >= a.in.H) continue;
    for (int kx = 0; kx < a.k; ++kx) {
      const int ix = ox * a.stride + kx - pad;
      if (ix < 0 || ix >= a.in.W) continue;
      const T* ip = in + ((size_t)iy * a.in.W + ix) * a.in.pix_stride;
      const int tap = ky * a.k + kx;
      for (int cv = 0; cv < cin_vec; ++cv) {
        float x[8];
        unpack8<T>(*reinterpret_cast<const uint4*>(ip + cv * 8), x);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float wv[8];
          unpack8<T>(*reinterpret_cast<const uint4*>(w + ((size_t)(cg * 8 + c) * taps + tap) * a.cin_pad + cv * 8), wv);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[c] = fmaf(x[i], wv[i], acc[c]);
        }
      }
    }
  }
  if (a.pow11_ch0 && cg == 0) acc[0] = powf(1.1f, acc[0]);
  if (a.has_res) {
    const T* rp = (const T*)a.res.ptr + ((size_t)b * a.res.H * a.res.W + (size_t)oy * Wo + ox) * a.res.pix_stride + cg * 8;
    float r[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(rp), r);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] += r[c];
  }
  if (a.relu) {
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = fmaxf(acc[c], 0.f);
  }
  const size_t opix = ((size_t)b * Ho + oy) * Wo + ox;
  if (a.out.dtype == ACR_DT_F32) {
    float* o = (float*)a.out.ptr + opix * a.out.pix_stride + cg * 8;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else {
    T* o = (T*)a.out.ptr + opix * a.out.pix_stride + cg * 8;
    *reinterpret_cast<uint4*>(o) = pack8<T>(acc);
  }
}

int launch_conv_ref(const ConvArgs& a, int act_dtype, cudaStream_t st) {
  ACR_CHECK_ARG(a.cout_pad % 8 == 0 && a.cin_pad % 8 == 0 && a.in.pix_stride % 8 == 0, "conv_ref: alignment");
  ACR_CHECK_ARG(!a.s2x && a.n_ext == 0, "conv_ref: x-paired stride-2 inputs / folded fuse sums exist on the tensor-core path only");
  const long long total = (long long)a.batch * a.out.H * a.out.W * (a.cout_pad / 8);
  ACR_DISPATCH_ACT(act_dtype, conv_ref_kernel<T><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, total));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------------ fuse
// HighResolutionModule.forward :677-684: y = relu(sum_j f_ij(x_j)), nearest upsample for j > i.
// thread = (pixel, 8 channels); fp32 sum in the reference's order, one rounding.
template <typename T, int CPT>   // CPT = channels per thread: 16 -> 256-bit accesses (whole sectors), 8 -> 128-bit
__global__ void __launch_bounds__(256) fuse_kernel(FuseArgs a) {
  // grid = (ceil(W * C/CPT / 256), H, B): one 32-bit division per thread, none on 64-bit values
  const unsigned ngrp = a.out.C / CPT;
  const unsigned tid = blockIdx.x * 256 + threadIdx.x;
  const unsigned x = tid / ngrp, cg = tid - x * ngrp;
  if (x >= (unsigned)a.out.W) return;
  const unsigned y = blockIdx.y;
  const size_t b = blockIdx.z;
  float acc[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) acc[c] = 0.f;
#pragma unroll 4
  for (int i = 0; i < a.n_in; ++i) {
    const TensorRef& t = a.in[i];
    const unsigned sx = x >> a.shift[i], sy = y >> a.shift[i];
    const T* p = (const T*)t.ptr + ((b * t.H + sy) * (size_t)t.W + sx) * t.pix_stride + cg * CPT;
    float v[CPT];
    if (CPT == 16) {
      uint4 u0, u1;
      ldg256(p, u0, u1);
      unpack8<T>(u0, v);
      unpack8<T>(u1, v + 8);
    } else {
      unpack8<T>(*reinterpret_cast<const uint4*>(p), v);
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = (i == 0) ? v[c] : acc[c] + v[c];
  }
  if (a.relu) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = fmaxf(acc[c], 0.f);
  }
  T* o = (T*)a.out.ptr + ((b * a.out.H + y) * (size_t)a.out.W + x) * a.out.pix_stride + cg * CPT;
  if (CPT == 16) stg256(o, pack8<T>(acc), pack8<T>(acc + 8));
  else *reinterpret_cast<uint4*>(o) = pack8<T>(acc);
}

static inline dim3 row_grid(const TensorRef& out, int batch, int cpt) {
  return dim3((unsigned)((out.W * (out.C / cpt) + 255) / 256), (unsigned)out.H, (unsigned)batch);
}
static inline bool rows32(const TensorRef& t) { return (uintptr_t)t.ptr % 32 == 0 && t.pix_stride % 16 == 0 && t.C % 16 == 0; }

int launch_fuse(const FuseArgs& a, int act_dtype, cudaStream_t st) {
  ACR_CHECK_ARG(a.n_in >= 1 && a.n_in <= 4 && a.out.C % 8 == 0 && a.out.H <= 65535 && a.batch <= 65535, "fuse: bad arguments");
  bool wide = rows32(a.out);
  for (int i = 0; i < a.n_in; ++i) {
    ACR_CHECK_ARG(a.in[i].C == a.out.C && (a.in[i].H << a.shift[i]) == a.out.H, "fuse: term %d shape mismatch", i);
    wide = wide && rows32(a.in[i]);
  }
  if (wide) ACR_DISPATCH_ACT(act_dtype, fuse_kernel<T, 16><<<row_grid(a.out, a.batch, 16), 256, 0, st>>>(a));
  else ACR_DISPATCH_ACT(act_dtype, fuse_kernel<T, 8><<<row_grid(a.out, a.batch, 8), 256, 0, st>>>(a));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------- bilinear x2
// Up.forward :432  F.interpolate(scale 2, bilinear, align_corners=True)
template <typename T, int CPT>
__global__ void __launch_bounds__(256) bilinear2x_kernel(TensorRef in, TensorRef out) {
  const unsigned ngrp = out.C / CPT;
  const unsigned tid = blockIdx.x * 256 + threadIdx.x;
  const unsigned xu = tid / ngrp, cg = tid - xu * ngrp;
  if (xu >= (unsigned)out.W) return;
  const int x = (int)xu, y = (int)blockIdx.y;
  const size_t b = blockIdx.z;
  const float sy = (float)(in.H - 1) / (float)(out.H - 1), sx = (float)(in.W - 1) / (float)(out.W - 1);
  const float fy = sy * y, fx = sx * x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, in.H - 1), x1 = min(x0 + 1, in.W - 1);
  const float ly = fy - y0, lx = fx - x0;
  const T* base = (const T*)in.ptr + b * in.img_stride() + cg * CPT;
  float v00[CPT], v01[CPT], v10[CPT], v11[CPT], o[CPT];
  auto load = [&](int yy, int xx, float* v) {
    const T* p = base + ((size_t)yy * in.W + xx) * in.pix_stride;
    if (CPT == 16) {
      uint4 u0, u1;
      ldg256(p, u0, u1);
      unpack8<T>(u0, v);
      unpack8<T>(u1, v + 8);
    } else {
      unpack8<T>(*reinterpret_cast<const uint4*>(p), v);
    }
  };
  load(y0, x0, v00); load(y0, x1, v01); load(y1, x0, v10); load(y1, x1, v11);
#pragma unroll
  for (int c = 0; c < CPT; ++c)
    o[c] = (1.f - ly) * ((1.f - lx) * v00[c] + lx * v01[c]) + ly * ((1.f - lx) * v10[c] + lx * v11[c]);
  T* op = (T*)out.ptr + ((b * out.H + y) * (size_t)out.W + x) * out.pix_stride + cg * CPT;
  if (CPT == 16) stg256(op, pack8<T>(o), pack8<T>(o + 8));
  else *reinterpret_cast<uint4*>(op) = pack8<T>(o);
}

int launch_bilinear2x(const TensorRef& in, const TensorRef& out, int batch, int act_dtype, cudaStream_t st) {
  ACR_CHECK_ARG(out.H == 2 * in.H && out.W == 2 * in.W && out.C == in.C && in.C % 8 == 0 && out.H <= 65535 && batch <= 65535,
                "bilinear2x: shapes");
  if (rows32(in) && rows32(out)) ACR_DISPATCH_ACT(act_dtype, bilinear2x_kernel<T, 16><<<row_grid(out, batch, 16), 256, 0, st>>>(in, out));
  else ACR_DISPATCH_ACT(act_dtype, bilinear2x_kernel<T, 8><<<row_grid(out, batch, 8), 256, 0, st>>>(in, out));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------------ coord
// get_coord_maps :340-369 + the cat at :52 -- channel c_off = x in [-1,1], c_off+1 = y, rest of the
// pad group zero.  thread = pixel, the 8-channel groups leave as 16-byte stores (whole sectors).
template <typename T>
__global__ void __launch_bounds__(256) coord_kernel(TensorRef out, int c_off, int npad) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= out.W) return;
  const size_t b = blockIdx.z;
  T* o = (T*)out.ptr + ((b * out.H + y) * (size_t)out.W + x) * out.pix_stride + c_off;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = 0.f;
  v[0] = (float)x / (float)(out.W - 1) * 2.f - 1.f;
  v[1] = (float)y / (float)(out.H - 1) * 2.f - 1.f;
  *reinterpret_cast<uint4*>(o) = pack8<T>(v);
  v[0] = v[1] = 0.f;
  for (int c = 8; c < npad; c += 8) *reinterpret_cast<uint4*>(o + c) = pack8<T>(v);
}

int launch_coord(const TensorRef& out, int c_off, int batch, int act_dtype, cudaStream_t st) {
  const int npad = out.pix_stride - c_off;
  ACR_CHECK_ARG(npad >= 8 && npad % 8 == 0 && c_off % 8 == 0 && out.pix_stride % 8 == 0 && out.H <= 65535 && batch <= 65535,
                "coord: the coord channels need an aligned 8-channel group");
  const dim3 grid((unsigned)((out.W + 255) / 256), (unsigned)out.H, (unsigned)batch);
  ACR_DISPATCH_ACT(act_dtype, coord_kernel<T><<<grid, 256, 0, st>>>(out, c_off, npad));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------- attention pooling
// Hadamard_product :103-113 on the contact features with part_attention = nearest-1/2 of the
// segmentation logits minus the background channel (:126-128).  Split-softmax: CTA (b, chunk) handles
// HW/POOL_CHUNKS pixels and emits the un-normalised sums acc[c][j] = sum_p exp(l_jp - m_j) f_pc together
// with (m_j, s_j); launch_parthead merges the chunks.
//
// The contraction over pixels is a [32 parts] x [256 channels] x [K = pixels] GEMM per image: warp-level
// tensor-core MMAs (m16n8k16, fp32 accumulate) keep the kernel on its HBM roofline (one streaming read of
// the feature map); the weights w = exp(l - m) are rounded to the storage type T once and s_j sums the
// ROUNDED values, so the normalised weights still sum to one.
constexpr int POOL_HALF = 512;                // pixels whose softmax weights are resident in shared memory
constexpr int POOL_W_STRIDE = POOL_HALF + 8;  // elements; +16 B keeps ldmatrix rows on distinct banks
constexpr int POOL_F_STRIDE = 40;             // elements; 32 channels + 16 B pad
constexpr int POOL_F_TILE = 32;               // pixels per cp.async stage
constexpr int POOL_STAGES = 3;
// 92.5 KB: two CTAs per SM, so one CTA's weight pass overlaps the other's feature stream
constexpr int POOL_SMEM_BYTES = 32 * POOL_W_STRIDE * 2 + 8 * POOL_STAGES * POOL_F_TILE * POOL_F_STRIDE * 2;

template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gmem) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(256, 2) pool_kernel(TensorRef feat, TensorRef logits, float* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char pool_smem[];
  T* s_w = reinterpret_cast<T*>(pool_smem);                                   // [32 parts][POOL_W_STRIDE]
  T* s_f = s_w + 32 * POOL_W_STRIDE;                                          // [8 warps][stages][tile][40]
  __shared__ float s_red[8][32];
  __shared__ float s_m[32], s_scale[32];
  const int b = blockIdx.x, chunk = blockIdx.y, t = threadIdx.x;
  const int HW = feat.H * feat.W, per = HW / POOL_CHUNKS, p0 = chunk * per;
  const T* lg = (const T*)logits.ptr + (size_t)b * logits.img_stride();
  const T* ft = (const T*)feat.ptr + (size_t)b * feat.img_stride();
  const int j = t & 31, warp = t >> 5, lane = j;
  if (t < 32) s_m[t] = -INFINITY;

  // this warp's feature stream: channels [32 warp, 32 warp + 32), POOL_F_TILE pixels per stage
  const int ntiles = per / POOL_F_TILE, tiles_per_half = POOL_HALF / POOL_F_TILE;
  T* my_f = s_f + (size_t)warp * POOL_STAGES * POOL_F_TILE * POOL_F_STRIDE;
  auto issue = [&](int tile) {
    if (tile < ntiles) {
      T* dst = my_f + (size_t)(tile % POOL_STAGES) * POOL_F_TILE * POOL_F_STRIDE;
      const T* src = ft + (size_t)(p0 + tile * POOL_F_TILE) * feat.pix_stride + warp * 32;
#pragma unroll
      for (int i = 0; i < POOL_F_TILE / 8; ++i) {
        const int id = i * 32 + lane, row = id >> 2, col = id & 3;
        cp_async16(dst + row * POOL_F_STRIDE + col * 8, src + (size_t)row * feat.pix_stride + col * 8);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
  issue(1);

  float acc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  const int frag_row = (lane & 7) + ((lane >> 3) & 1) * 8, frag_col = (lane >> 4) * 8;
  const int g = lane >> 2, tq = lane & 3;
  float ssum = 0.f;   // this thread's share of sum_p w[j][p], relative to the running maximum s_m[j]

  for (int h = 0; h < per / POOL_HALF; ++h) {
    if (h > 0) __syncthreads();   // every warp is done with the previous half's weights and scales
    // ---- softmax weights of this half: raw logits -> running maximum -> exp, rounded to T, [part][pixel]
    float mloc = -INFINITY;
#pragma unroll 8
    for (int p = warp; p < POOL_HALF; p += 8) {
      const int P = p0 + h * POOL_HALF + p, y = P / feat.W, x = P % feat.W;
      const T v = lg[((size_t)(2 * y) * logits.W + 2 * x) * logits.pix_stride + 1 + j];
      s_w[j * POOL_W_STRIDE + p] = v;
      mloc = fmaxf(mloc, to_f32<T>(v));
    }
    s_red[warp][j] = mloc;
    __syncthreads();
    if (t < 32) {
      float mm = s_red[0][t];
#pragma unroll
      for (int i = 1; i < 8; ++i) mm = fmaxf(mm, s_red[i][t]);
      const float m_old = s_m[t], m_new = fmaxf(m_old, mm);
      s_scale[t] = expf(m_old - m_new);   // 0 for the first half
      s_m[t] = m_new;
    }
    __syncthreads();
    const float m = s_m[j];
    ssum *= s_scale[j];
    for (int p = warp; p < POOL_HALF; p += 8) {
      const T w = from_f32<T>(expf(to_f32<T>(s_w[j * POOL_W_STRIDE + p]) - m));
      s_w[j * POOL_W_STRIDE + p] = w;
      ssum += to_f32<T>(w);
    }
    __syncthreads();
    if (h > 0) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float f0 = s_scale[mt * 16 + g], f1 = s_scale[mt * 16 + g + 8];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          acc[mt][nt][0] *= f0; acc[mt][nt][1] *= f0; acc[mt][nt][2] *= f1; acc[mt][nt][3] *= f1;
        }
      }
    }
    // ---- acc[part][channel] += w[part][pixel] * f[pixel][channel]
    for (int lt = 0; lt < tiles_per_half; ++lt) {
      const int tile = h * tiles_per_half + lt;
      issue(tile + 2);
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncwarp();
      const T* sf = my_f + (size_t)(tile % POOL_STAGES) * POOL_F_TILE * POOL_F_STRIDE;
#pragma unroll
      for (int ks = 0; ks < POOL_F_TILE / 16; ++ks) {
        uint32_t a[2][4], bq[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
          ldsm_x4(a[mt], s_w + (size_t)(mt * 16 + frag_row) * POOL_W_STRIDE + lt * POOL_F_TILE + ks * 16 + frag_col);
#pragma unroll
        for (int np = 0; np < 2; ++np)
          ldsm_x4_trans(bq[np], sf + (size_t)(ks * 16 + frag_row) * POOL_F_STRIDE + np * 16 + frag_col);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            mma16816<T>(acc[mt][nt], a[mt], bq[nt >> 1][(nt & 1) * 2], bq[nt >> 1][(nt & 1) * 2 + 1]);
      }
      __syncwarp();
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  s_red[warp][j] = ssum;
  __syncthreads();

  float* o = part + ((size_t)b * POOL_CHUNKS + chunk) * POOL_PART_FLOATS;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = warp * 32 + nt * 8 + 2 * tq, jj = mt * 16 + g;
      o[(size_t)c * 32 + jj] = acc[mt][nt][0];
      o[(size_t)(c + 1) * 32 + jj] = acc[mt][nt][1];
      o[(size_t)c * 32 + jj + 8] = acc[mt][nt][2];
      o[(size_t)(c + 1) * 32 + jj + 8] = acc[mt][nt][3];
    }
  if (t < 32) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += s_red[i][t];
    o[256 * 32 + t] = s_m[t];
    o[256 * 32 + 32 + t] = sum;
  }
}

int launch_pool(const TensorRef& feat, const TensorRef& logits, float* part, int batch, int act_dtype,
                cudaStream_t st) {
  if (pool_tc_enabled() && (act_dtype == ACR_DT_BF16 || act_dtype == ACR_DT_F16) && logits.H == 2 * feat.H && logits.W == 2 * feat.W &&
      logits.C >= 33 && pool_tc_takes(feat, logits))
    return launch_pool_tc(feat, logits, part, batch, act_dtype, st);     // tcgen05 / TMA form (pool_tc.cu)
  const int per = feat.H * feat.W / POOL_CHUNKS;
  ACR_CHECK_ARG(feat.C == 256 && logits.H == 2 * feat.H && logits.W == 2 * feat.W && logits.C >= 33 &&
                    (feat.H * feat.W) % POOL_CHUNKS == 0 && per % POOL_HALF == 0 &&
                    feat.pix_stride % 8 == 0, "pool: shapes");
  ACR_DISPATCH_ACT(act_dtype, {
    static unsigned long long attr_set = 0;
    ACR_CHECK_CUDA(ensure_dynamic_smem(pool_kernel<T>, POOL_SMEM_BYTES, &attr_set));
    pool_kernel<T><<<dim3(batch, POOL_CHUNKS), 256, POOL_SMEM_BYTES, st>>>(feat, logits, part);
  });
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// --------------------------------------------------------------------------------- part head
// part_forward :139-164 after the pooling: LocallyConnected2d (:559-569) on the 16 parts of each
// hand (left = parts 16..31, right = 0..15), the 256->64 1x1 conv applied to the pooled feature
// (softmax weights sum to 1, so conv-then-pool == pool-then-conv), Linear 1024->10, and the
// spatially-constant half of the 218->109 conv folded into a per-image bias.  One CTA per image.
__global__ void __launch_bounds__(256) parthead_kernel(PartHeadArgs a) {
  __shared__ float s_pool[256][33];
  __shared__ float s_sf[64][33];
  __shared__ float s_scale[POOL_CHUNKS][32];
  __shared__ float s_pare[2][112];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* part = a.part + (size_t)b * POOL_CHUNKS * POOL_PART_FLOATS;
  if (t < 32) {
    float M = -INFINITY;
    for (int c = 0; c < POOL_CHUNKS; ++c) M = fmaxf(M, part[(size_t)c * POOL_PART_FLOATS + 256 * 32 + t]);
    float S = 0.f;
    for (int c = 0; c < POOL_CHUNKS; ++c) {
      const float e = expf(part[(size_t)c * POOL_PART_FLOATS + 256 * 32 + t] - M);
      s_scale[c][t] = e;
      S += e * part[(size_t)c * POOL_PART_FLOATS + 256 * 32 + 32 + t];
    }
    for (int c = 0; c < POOL_CHUNKS; ++c) s_scale[c][t] /= S;
  }
  __syncthreads();
  {  // merge the chunks: element e = channel * 32 + part of the [256][32] partial sums; thread t takes e = t + 256 i, so a warp
     // reads 128 contiguous bytes per load and its part index (t & 31) never changes (same chunk order as before: same bits)
    const int j = t & 31;
    float sc[POOL_CHUNKS];
#pragma unroll
    for (int c = 0; c < POOL_CHUNKS; ++c) sc[c] = s_scale[c][j];
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const int e = t + 256 * i;
      float v = 0.f;
#pragma unroll
      for (int c = 0; c < POOL_CHUNKS; ++c) v = fmaf(__ldg(part + (size_t)c * POOL_PART_FLOATS + e), sc[c], v);
      s_pool[e >> 5][j] = v;
      a.pooled[(size_t)b * 256 * 32 + e] = v;
    }
  }
  __syncthreads();
  // shape features: sf[c64][j] = W(64,256) . pooled[:, j] + b
  {  // thread t: part j = t & 31, output channels c0 + 8 i -- one shared-memory read of pooled[c][j] serves eight outputs; every
     // output still accumulates over c in ascending order (same bits as one output at a time)
    const int j = t & 31, c0 = t >> 5;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = a.shape_b[c0 + 8 * i];
    const float* w = a.shape_w + (size_t)c0 * 256;
#pragma unroll 4
    for (int c = 0; c < 256; ++c) {
      const float x = s_pool[c][j];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(__ldg(w + (size_t)i * 8 * 256 + c), x, v[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_sf[c0 + 8 * i][j] = v[i];
  }
  // contact offsets: thread (side, joint, o)
  if (t < 192) {
    const int side = t / 96, r = t % 96, jj = r / 6, o = r % 6;
    const int pj = side == 0 ? 16 + jj : jj;
    const float* w = a.lc_w[side] + (size_t)o * 256 * 16 + jj;   // (6,256,16)
    float v = 0.f;
    for (int c = 0; c < 256; ++c) v = fmaf(s_pool[c][pj], w[(size_t)c * 16], v);
    s_pare[side][jj * 6 + o] = v;
  }
  __syncthreads();
  // shape offsets: Linear(1024 -> 10) on flatten(sf[:, parts]) (c64-major), one warp per output
  {
    const int warp = t >> 5, lane = t & 31;
    for (int oi = warp; oi < 20; oi += 8) {
      const int side = oi / 10, o = oi % 10;
      const float* w = a.lin_w[side] + (size_t)o * 1024;
      float v = 0.f;
      for (int i = lane; i < 1024; i += 32) {
        const int c64 = i >> 4, jj = i & 15;
        v = fmaf(w[i], s_sf[c64][side == 0 ? 16 + jj : jj], v);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) s_pare[side][96 + o] = v + a.lin_b[side][o];
    }
  }
  __syncthreads();
  // per-image bias of the folded final conv: b + W[:, 112:218] . pare
  for (int oi = t; oi < 2 * 112; oi += 256) {
    const int side = oi / 112, o = oi % 112;
    float v = 0.f;
    if (o < 109) {
      v = a.fin_b[side][o];
      const float* w = a.fin_w[side] + (size_t)o * 218 + 112;
      for (int i = 0; i < 106; ++i) v = fmaf(w[i], s_pare[side][i], v);
    }
    a.bias_img[side][(size_t)b * 112 + o] = v;
    if (o < 106) a.pare[side][(size_t)b * 106 + o] = s_pare[side][o];
  }
}

int launch_parthead(const PartHeadArgs& a, cudaStream_t st) {
  parthead_kernel<<<a.batch, 256, 0, st>>>(a);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

}  // namespace acr
