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
                                                          int H, int W, int out_stride, long long total) {
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= total) return;
  const int Ho = H / 2, Wo = W / 2;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho);
  const long long b = pix / ((long long)Wo * Ho);
  float v[32];
#pragma unroll
  for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - 1;
    const bool yok = iy >= 0 && iy < H;
    const uint8_t* row = img + ((b * H + (yok ? iy : 0)) * (long long)W) * 3;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - 1;
      const bool ok = yok && ix >= 0 && ix < W;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        v[(ky * 3 + kx) * 3 + ci] = ok ? (float)row[ix * 3 + ci] / 255.f * 2.0f - 1.0f : 0.f;
    }
  }
  T* o = out + pix * out_stride;
#pragma unroll
  for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(o + c * 8) = pack8<T>(v + c * 8);
}

int launch_im2col_stem(const TensorRef& img, const TensorRef& out, int batch, int act_dtype, cudaStream_t st) {
  ACR_CHECK_ARG(out.C == 32 && out.H * 2 == img.H && out.W * 2 == img.W && img.dtype == ACR_DT_U8 && out.pix_stride >= 32,
                "im2col_stem: shape mismatch");
  const long long total = (long long)batch * out.H * out.W;
  ACR_DISPATCH_ACT(act_dtype, im2col_stem_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                                  (const uint8_t*)img.ptr, (T*)out.ptr, img.H, img.W, out.pix_stride, total));
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
  const long long total = (long long)a.batch * a.out.H * a.out.W * (a.cout_pad / 8);
  ACR_DISPATCH_ACT(act_dtype, conv_ref_kernel<T><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, total));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------------ fuse
// HighResolutionModule.forward :677-684: y = relu(sum_j f_ij(x_j)), nearest upsample for j > i.
// thread = (pixel, 8 channels); fp32 sum in the reference's order, one rounding.
template <typename T>
__global__ void __launch_bounds__(256) fuse_kernel(FuseArgs a, long long total) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gid >= total) return;
  const int ngrp = a.out.C / 8;
  const int cg = (int)(gid % ngrp);
  const long long pix = gid / ngrp;
  const int W = a.out.W, H = a.out.H;
  const int x = (int)(pix % W), y = (int)((pix / W) % H);
  const size_t b = (size_t)(pix / ((long long)W * H));
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  for (int i = 0; i < a.n_in; ++i) {
    const TensorRef& t = a.in[i];
    const int sx = x >> a.shift[i], sy = y >> a.shift[i];
    const T* p = (const T*)t.ptr + (b * t.H * t.W + (size_t)sy * t.W + sx) * t.pix_stride + cg * 8;
    float v[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(p), v);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = (i == 0) ? v[c] : acc[c] + v[c];
  }
  if (a.relu) {
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = fmaxf(acc[c], 0.f);
  }
  T* o = (T*)a.out.ptr + (size_t)pix * a.out.pix_stride + cg * 8;
  *reinterpret_cast<uint4*>(o) = pack8<T>(acc);
}

int launch_fuse(const FuseArgs& a, int act_dtype, cudaStream_t st) {
  ACR_CHECK_ARG(a.n_in >= 1 && a.n_in <= 4 && a.out.C % 8 == 0, "fuse: bad arguments");
  for (int i = 0; i < a.n_in; ++i)
    ACR_CHECK_ARG(a.in[i].C == a.out.C && (a.in[i].H << a.shift[i]) == a.out.H, "fuse: term %d shape mismatch", i);
  const long long total = (long long)a.batch * a.out.H * a.out.W * (a.out.C / 8);
  ACR_DISPATCH_ACT(act_dtype, fuse_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, total));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------- bilinear x2
// Up.forward :432  F.interpolate(scale 2, bilinear, align_corners=True)
template <typename T>
__global__ void __launch_bounds__(256) bilinear2x_kernel(TensorRef in, TensorRef out, long long total) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gid >= total) return;
  const int ngrp = out.C / 8;
  const int cg = (int)(gid % ngrp);
  const long long pix = gid / ngrp;
  const int x = (int)(pix % out.W), y = (int)((pix / out.W) % out.H);
  const size_t b = (size_t)(pix / ((long long)out.W * out.H));
  const float sy = (float)(in.H - 1) / (float)(out.H - 1), sx = (float)(in.W - 1) / (float)(out.W - 1);
  const float fy = sy * y, fx = sx * x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, in.H - 1), x1 = min(x0 + 1, in.W - 1);
  const float ly = fy - y0, lx = fx - x0;
  const T* base = (const T*)in.ptr + b * in.img_stride() + cg * 8;
  float v00[8], v01[8], v10[8], v11[8], o[8];
  unpack8<T>(*reinterpret_cast<const uint4*>(base + ((size_t)y0 * in.W + x0) * in.pix_stride), v00);
  unpack8<T>(*reinterpret_cast<const uint4*>(base + ((size_t)y0 * in.W + x1) * in.pix_stride), v01);
  unpack8<T>(*reinterpret_cast<const uint4*>(base + ((size_t)y1 * in.W + x0) * in.pix_stride), v10);
  unpack8<T>(*reinterpret_cast<const uint4*>(base + ((size_t)y1 * in.W + x1) * in.pix_stride), v11);
#pragma unroll
  for (int c = 0; c < 8; ++c)
    o[c] = (1.f - ly) * ((1.f - lx) * v00[c] + lx * v01[c]) + ly * ((1.f - lx) * v10[c] + lx * v11[c]);
  *reinterpret_cast<uint4*>((T*)out.ptr + (size_t)pix * out.pix_stride + cg * 8) = pack8<T>(o);
}

int launch_bilinear2x(const TensorRef& in, const TensorRef& out, int batch, int act_dtype, cudaStream_t st) {
  ACR_CHECK_ARG(out.H == 2 * in.H && out.W == 2 * in.W && out.C == in.C && in.C % 8 == 0, "bilinear2x: shapes");
  const long long total = (long long)batch * out.H * out.W * (out.C / 8);
  ACR_DISPATCH_ACT(act_dtype, bilinear2x_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, total));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------------ coord
// get_coord_maps :340-369 + the cat at :52 -- channel c_off = x in [-1,1], c_off+1 = y, rest of the
// 16-channel pad group zero.  thread = pixel.
template <typename T>
__global__ void __launch_bounds__(256) coord_kernel(TensorRef out, int c_off, int npad, long long total) {
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= total) return;
  const int x = (int)(pix % out.W), y = (int)((pix / out.W) % out.H);
  T* o = (T*)out.ptr + (size_t)pix * out.pix_stride + c_off;
  o[0] = from_f32<T>((float)x / (float)(out.W - 1) * 2.f - 1.f);
  o[1] = from_f32<T>((float)y / (float)(out.H - 1) * 2.f - 1.f);
  for (int c = 2; c < npad; ++c) o[c] = from_f32<T>(0.f);
}

int launch_coord(const TensorRef& out, int c_off, int batch, int act_dtype, cudaStream_t st) {
  const int npad = out.pix_stride - c_off;
  ACR_CHECK_ARG(npad >= 2, "coord: no room for the coord channels");
  const long long total = (long long)batch * out.H * out.W;
  ACR_DISPATCH_ACT(act_dtype, coord_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, c_off, npad, total));
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------- attention pooling
// Hadamard_product :103-113 on the contact features with part_attention = nearest-1/2 of the
// segmentation logits minus the background channel (:126-128).  Split-softmax: CTA (b, chunk)
// handles 1024 pixels, emits un-normalised sums acc[c][j] = sum_p exp(l_jp - m_j) f_pc together
// with (m_j, s_j); launch_parthead merges the 16 chunks.  thread = feature channel.
template <typename T>
__global__ void __launch_bounds__(256) pool_kernel(TensorRef feat, TensorRef logits, float* __restrict__ part) {
  __shared__ __align__(16) float s_w[32][32];   // [pixel][part]
  __shared__ float s_red[8][32];
  __shared__ float s_m[32];
  const int b = blockIdx.x, chunk = blockIdx.y, t = threadIdx.x;
  const int HW = feat.H * feat.W, per = HW / POOL_CHUNKS, p0 = chunk * per;
  const T* lg = (const T*)logits.ptr + (size_t)b * logits.img_stride();
  const T* ft = (const T*)feat.ptr + (size_t)b * feat.img_stride();
  const int j = t & 31, sub = t >> 5;  // this thread's part channel / pixel sub-lane
  auto logit = [&](int p) -> float {
    const int y = p / feat.W, x = p % feat.W;
    return to_f32<T>(lg[((size_t)(2 * y) * logits.W + 2 * x) * logits.pix_stride + 1 + j]);
  };
  // chunk maximum per part
  float m = -INFINITY;
  for (int p = p0 + sub; p < p0 + per; p += 8) m = fmaxf(m, logit(p));
  s_red[sub][j] = m;
  __syncthreads();
  if (t < 32) {
    float mm = s_red[0][t];
#pragma unroll
    for (int i = 1; i < 8; ++i) mm = fmaxf(mm, s_red[i][t]);
    s_m[t] = mm;
  }
  __syncthreads();
  m = s_m[j];
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  float ssum = 0.f;
  for (int q = p0; q < p0 + per; q += 32) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pp = sub + 8 * i;
      const float w = expf(logit(q + pp) - m);
      s_w[pp][j] = w;
      ssum += w;
    }
    __syncthreads();
    for (int p8 = 0; p8 < 32; p8 += 8) {
      T fv[8];   // 8 independent loads in flight before the FMA block
#pragma unroll
      for (int u = 0; u < 8; ++u) fv[u] = ft[(size_t)(q + p8 + u) * feat.pix_stride + t];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float f = to_f32<T>(fv[u]);
        const float4* wr = reinterpret_cast<const float4*>(&s_w[p8 + u][0]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w4 = wr[i];
          acc[i * 4 + 0] = fmaf(w4.x, f, acc[i * 4 + 0]); acc[i * 4 + 1] = fmaf(w4.y, f, acc[i * 4 + 1]);
          acc[i * 4 + 2] = fmaf(w4.z, f, acc[i * 4 + 2]); acc[i * 4 + 3] = fmaf(w4.w, f, acc[i * 4 + 3]);
        }
      }
    }
  }
  __syncthreads();
  s_red[sub][j] = ssum;
  __syncthreads();
  float* o = part + ((size_t)b * POOL_CHUNKS + chunk) * POOL_PART_FLOATS;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(o + t * 32 + i * 4) = make_float4(acc[i * 4], acc[i * 4 + 1], acc[i * 4 + 2], acc[i * 4 + 3]);
  if (t < 32) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += s_red[i][t];
    o[256 * 32 + t] = s_m[t];
    o[256 * 32 + 32 + t] = s;
  }
}

int launch_pool(const TensorRef& feat, const TensorRef& logits, float* part, int batch, int act_dtype,
                cudaStream_t st) {
  ACR_CHECK_ARG(feat.C == 256 && logits.H == 2 * feat.H && logits.W == 2 * feat.W && logits.C >= 33 &&
                    (feat.H * feat.W) % (POOL_CHUNKS * 32) == 0, "pool: shapes");
  ACR_DISPATCH_ACT(act_dtype, pool_kernel<T><<<dim3(batch, POOL_CHUNKS), 256, 0, st>>>(feat, logits, part));
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
  for (int j = 0; j < 32; ++j) {
    float v = 0.f;
    for (int c = 0; c < POOL_CHUNKS; ++c) v = fmaf(part[(size_t)c * POOL_PART_FLOATS + t * 32 + j], s_scale[c][j], v);
    s_pool[t][j] = v;
    a.pooled[((size_t)b * 256 + t) * 32 + j] = v;
  }
  __syncthreads();
  // shape features: sf[c64][j] = W(64,256) . pooled[:, j] + b
  for (int o = t; o < 64 * 32; o += 256) {
    const int c64 = o >> 5, j = o & 31;
    float v = a.shape_b[c64];
    const float* w = a.shape_w + (size_t)c64 * 256;
    for (int c = 0; c < 256; ++c) v = fmaf(w[c], s_pool[c][j], v);
    s_sf[c64][j] = v;
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
