// fp32-storage validation kernels: the same launch plan as the 16-bit product path (netspec -> engine ->
// plan.cu), but every activation and weight is fp32 and every contraction accumulates in fp64 on the CUDA
// cores.  This is what `model_precision = 'fp32'` (the reference's shipped default, /root/reference/
// configs/demo.yml:7, acr/config.py:96, acr/model.py:33-41) maps to: it exists to pin the WHOLE pipeline
// (maps -> parse -> MANO -> vertices) against the reference's fp32 goldens at the 1e-4 the north star
// asks for, which 16-bit storage cannot show.  Speed is not a goal here (thread = pixel x 8 channels).
// Reference call sites are the same as in elementwise.cu; semantics are identical op for op.
#include "ops.cuh"

namespace acr {

// ------------------------------------------------------------------------------------ stem
// HigherResolutionNet.forward :832-835: x/255*2-1, conv1 3x3 s2 (3->64) + bn1 + relu.
// w fp32 [27][64] (tap-major, BN folded), bias fp32 [64]; thread = (pixel, channel).
__global__ void __launch_bounds__(256) stem_f32_kernel(const uint8_t* __restrict__ img, float* __restrict__ out,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       int H, int W, int out_stride, long long total) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gid >= total) return;
  const int c = (int)(gid & 63);
  const long long pix = gid >> 6;
  const int Ho = H / 2, Wo = W / 2;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho);
  const long long b = pix / ((long long)Wo * Ho);
  double acc = bias[c];
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - 1;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - 1;
      if (ix < 0 || ix >= W) continue;
      const uint8_t* px = img + ((b * H + iy) * W + ix) * 3;
      for (int ci = 0; ci < 3; ++ci) {
        const float xn = (float)px[ci] / 255.f * 2.0f - 1.0f;   // the reference's fp32 normalisation
        acc += (double)xn * (double)w[((ky * 3 + kx) * 3 + ci) * 64 + c];
      }
    }
  }
  out[pix * out_stride + c] = fmaxf((float)acc, 0.f);
}

int launch_stem_f32(const TensorRef& img, const TensorRef& out, const float* w, const float* bias, int batch,
                    cudaStream_t st) {
  ACR_CHECK_ARG(out.C == 64 && out.H * 2 == img.H && out.W * 2 == img.W && img.dtype == ACR_DT_U8 && out.dtype == ACR_DT_F32,
                "stem_f32: shape mismatch");
  const long long total = (long long)batch * out.H * out.W * 64;
  stem_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint8_t*)img.ptr, (float*)out.ptr, w, bias, img.H,
                                                                   img.W, out.pix_stride, total);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------------ conv
// Same contract as the tcgen05 conv (ConvArgs) with fp32 tensors: weights [cout_pad][k*k][cin_pad] fp32,
// fp64 accumulate, epilogue = +bias (+residual) (ReLU) (1.1**x on channel 0), fp32 NHWC output.
__global__ void __launch_bounds__(128) conv_f32_kernel(ConvArgs a, long long total) {
  const long long gid = (long long)blockIdx.x * 128 + threadIdx.x;
  if (gid >= total) return;
  const int ngrp = a.cout_pad / 8;
  const int cg = (int)(gid % ngrp);
  const long long pix = gid / ngrp;
  const int Wo = a.out.W, Ho = a.out.H;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho);
  const int b = (int)(pix / ((long long)Wo * Ho));
  const float* in = (const float*)a.in.ptr + (size_t)b * a.in.img_stride();
  const float* w = (const float*)a.w;
  const int pad = a.k / 2, taps = a.k * a.k;
  double acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.0;
  const int cin_vec = (a.in.C + 3) / 4;   // channels physically present (buffers are padded to 16)
  for (int ky = 0; ky < a.k; ++ky) {
    const int iy = oy * a.stride + ky - pad;
    if (iy < 0 || iy >= a.in.H) continue;
    for (int kx = 0; kx < a.k; ++kx) {
      const int ix = ox * a.stride + kx - pad;
      if (ix < 0 || ix >= a.in.W) continue;
      const float* ip = in + ((size_t)iy * a.in.W + ix) * a.in.pix_stride;
      const int tap = ky * a.k + kx;
      for (int cv = 0; cv < cin_vec; ++cv) {
        const float4 x = *reinterpret_cast<const float4*>(ip + cv * 4);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 wv = *reinterpret_cast<const float4*>(w + ((size_t)(cg * 8 + c) * taps + tap) * a.cin_pad + cv * 4);
          acc[c] += (double)x.x * wv.x + (double)x.y * wv.y + (double)x.z * wv.z + (double)x.w * wv.w;
        }
      }
    }
  }
  const float* bias = a.bias + (a.bias_per_image ? (size_t)b * a.cout_pad : 0) + cg * 8;
  float o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = (float)(acc[c] + (double)bias[c]);
  if (a.pow11_ch0 && cg == 0) o[0] = powf(1.1f, o[0]);
  if (a.has_res) {
    const float* rp = (const float*)a.res.ptr + ((size_t)b * a.res.H * a.res.W + (size_t)oy * Wo + ox) * a.res.pix_stride + cg * 8;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] += rp[c];
  }
  if (a.relu) {
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = fmaxf(o[c], 0.f);
  }
  float* op = (float*)a.out.ptr + (((size_t)b * Ho + oy) * Wo + ox) * a.out.pix_stride + cg * 8;
  *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
}

int launch_conv_f32(const ConvArgs& a, cudaStream_t st) {
  ACR_CHECK_ARG(a.cout_pad % 8 == 0 && a.cin_pad % 4 == 0 && a.in.pix_stride % 4 == 0 && a.in.dtype == ACR_DT_F32 &&
                    a.out.dtype == ACR_DT_F32 && !a.xpair && !a.s2x && a.n_ext == 0, "conv_f32: alignment / dtype / product-only forms");
  const long long total = (long long)a.batch * a.out.H * a.out.W * (a.cout_pad / 8);
  conv_f32_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, total);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------------ fuse
// HighResolutionModule.forward :677-684, fp32 sum in the reference's order (j = 0 .. nb-1).
__global__ void __launch_bounds__(256) fuse_f32_kernel(FuseArgs a, long long total) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gid >= total) return;
  const int C = a.out.C;
  const int c = (int)(gid % C);
  const long long pix = gid / C;
  const int x = (int)(pix % a.out.W), y = (int)((pix / a.out.W) % a.out.H);
  const size_t b = (size_t)(pix / ((long long)a.out.W * a.out.H));
  float acc = 0.f;
  for (int i = 0; i < a.n_in; ++i) {
    const TensorRef& t = a.in[i];
    const float v = ((const float*)t.ptr)[((b * t.H + (y >> a.shift[i])) * (size_t)t.W + (x >> a.shift[i])) * t.pix_stride + c];
    acc = (i == 0) ? v : acc + v;
  }
  if (a.relu) acc = fmaxf(acc, 0.f);
  ((float*)a.out.ptr)[((b * a.out.H + y) * (size_t)a.out.W + x) * a.out.pix_stride + c] = acc;
}

int launch_fuse_f32(const FuseArgs& a, cudaStream_t st) {
  ACR_CHECK_ARG(a.n_in >= 1 && a.n_in <= 4, "fuse_f32: bad arguments");
  for (int i = 0; i < a.n_in; ++i)
    ACR_CHECK_ARG(a.in[i].C == a.out.C && (a.in[i].H << a.shift[i]) == a.out.H, "fuse_f32: term %d shape mismatch", i);
  const long long total = (long long)a.batch * a.out.H * a.out.W * a.out.C;
  fuse_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, total);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------- bilinear x2
// Up.forward :432  F.interpolate(scale 2, bilinear, align_corners=True); ATen's fp32 formulation
// (source index = scale * dst, lambda = index - floor, weights (1-l), l).
__global__ void __launch_bounds__(256) bilinear2x_f32_kernel(TensorRef in, TensorRef out, long long total) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gid >= total) return;
  const int C = out.C;
  const int c = (int)(gid % C);
  const long long pix = gid / C;
  const int x = (int)(pix % out.W), y = (int)((pix / out.W) % out.H);
  const size_t b = (size_t)(pix / ((long long)out.W * out.H));
  const float sy = (float)(in.H - 1) / (float)(out.H - 1), sx = (float)(in.W - 1) / (float)(out.W - 1);
  const float fy = sy * y, fx = sx * x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, in.H - 1), x1 = min(x0 + 1, in.W - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float* base = (const float*)in.ptr + b * in.img_stride() + c;
  auto at = [&](int yy, int xx) { return base[((size_t)yy * in.W + xx) * in.pix_stride]; };
  const float v = (1.f - ly) * ((1.f - lx) * at(y0, x0) + lx * at(y0, x1)) + ly * ((1.f - lx) * at(y1, x0) + lx * at(y1, x1));
  ((float*)out.ptr)[((b * out.H + y) * (size_t)out.W + x) * out.pix_stride + c] = v;
}

int launch_bilinear2x_f32(const TensorRef& in, const TensorRef& out, int batch, cudaStream_t st) {
  ACR_CHECK_ARG(out.H == 2 * in.H && out.W == 2 * in.W && out.C == in.C, "bilinear2x_f32: shapes");
  const long long total = (long long)batch * out.H * out.W * out.C;
  bilinear2x_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, total);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------------------ coord
// get_coord_maps :340-369 + the cat at :52: channel c_off = x in [-1,1], c_off+1 = y, rest of the pad zero.
__global__ void __launch_bounds__(256) coord_f32_kernel(TensorRef out, int c_off, int npad, long long total) {
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= total) return;
  const int x = (int)(pix % out.W), y = (int)((pix / out.W) % out.H);
  float* o = (float*)out.ptr + pix * out.pix_stride + c_off;
  for (int c = 0; c < npad; ++c) o[c] = 0.f;
  o[0] = (float)x / (float)(out.W - 1) * 2.f - 1.f;
  o[1] = (float)y / (float)(out.H - 1) * 2.f - 1.f;
}

int launch_coord_f32(const TensorRef& out, int c_off, int batch, cudaStream_t st) {
  const int npad = out.pix_stride - c_off;
  ACR_CHECK_ARG(npad >= 2, "coord_f32: no room for the coord channels");
  const long long total = (long long)batch * out.H * out.W;
  coord_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, c_off, npad, total);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

// ------------------------------------------------------------------------- attention pooling
// Hadamard_product :103-113 with part_attention = nearest-1/2 of the segmentation logits minus the
// background channel (:126-128).  CTA = (image, part j): softmax over the HW pixels in fp32 like the
// reference, weighted channel sums in fp64.  Output in the partials layout of pool_kernel so that
// parthead_kernel is shared: chunk 0 carries the result, the other chunks are neutral (m = -inf, s = 0).
__global__ void __launch_bounds__(256) pool_f32_kernel(TensorRef feat, TensorRef logits, float* __restrict__ part) {
  __shared__ float s_red[256];
  __shared__ double s_acc[256];
  const int b = blockIdx.x, j = blockIdx.y, t = threadIdx.x;
  const int HW = feat.H * feat.W;
  const float* lg = (const float*)logits.ptr + (size_t)b * logits.img_stride() + 1 + j;
  const float* ft = (const float*)feat.ptr + (size_t)b * feat.img_stride();
  auto logit = [&](int P) {
    const int y = P / feat.W, x = P % feat.W;
    return lg[((size_t)(2 * y) * logits.W + 2 * x) * logits.pix_stride];
  };
  float m = -INFINITY;
  for (int P = t; P < HW; P += 256) m = fmaxf(m, logit(P));
  s_red[t] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) s_red[t] = fmaxf(s_red[t], s_red[t + o]);
    __syncthreads();
  }
  m = s_red[0];
  __syncthreads();
  double sum = 0.0;
  for (int P = t; P < HW; P += 256) sum += (double)expf(logit(P) - m);
  s_acc[t] = sum;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) s_acc[t] += s_acc[t + o];
    __syncthreads();
  }
  const double S = s_acc[0];
  // thread = channel t: acc = sum_p exp(l_p - m) f[p][t]
  double acc = 0.0;
  for (int P = 0; P < HW; ++P) acc += (double)expf(logit(P) - m) * (double)ft[(size_t)P * feat.pix_stride + t];
  float* o0 = part + (size_t)b * POOL_CHUNKS * POOL_PART_FLOATS;
  o0[(size_t)t * 32 + j] = (float)acc;
  if (t == 0) { o0[256 * 32 + j] = m; o0[256 * 32 + 32 + j] = (float)S; }
  for (int c = 1; c < POOL_CHUNKS; ++c) {
    float* oc = o0 + (size_t)c * POOL_PART_FLOATS;
    oc[(size_t)t * 32 + j] = 0.f;
    if (t == 0) { oc[256 * 32 + j] = -INFINITY; oc[256 * 32 + 32 + j] = 0.f; }
  }
}

int launch_pool_f32(const TensorRef& feat, const TensorRef& logits, float* part, int batch, cudaStream_t st) {
  ACR_CHECK_ARG(feat.C == 256 && logits.H == 2 * feat.H && logits.W == 2 * feat.W && logits.C >= 33 &&
                    feat.dtype == ACR_DT_F32 && logits.dtype == ACR_DT_F32, "pool_f32: shapes");
  pool_f32_kernel<<<dim3(batch, 32), 256, 0, st>>>(feat, logits, part);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

}  // namespace acr
