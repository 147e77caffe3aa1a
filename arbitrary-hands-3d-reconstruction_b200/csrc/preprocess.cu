// Frame pre-processing on the device: BGR -> RGB, white pad to a square, bicubic resize to
// input_size x input_size, all in one kernel reading the raw frame once.
// Replaces (reference, /root/reference/acr/utils.py): img_preprocess :1315-1337 (image[:,:,::-1],
// process_image_ori :1310-1313 -> image_pad_white_bg :1303-1308 with imgaug's
// compute_paddings_to_reach_aspect_ratio and pad_cval=255, cv2.resize(..., INTER_CUBIC) :1320).
//
// Arithmetic = OpenCV's generic 8-bit cubic path (third-party, absent from the reference tree: opencv-python,
// resize.cpp HResizeCubic/VResizeCubic with INTER_RESIZE_COEF_BITS = 11): per axis 4 taps with short
// coefficients round(c*2048) of the A=-0.75 cubic at fx = (d+0.5)*scale-0.5, border = replicate, horizontal
// pass exact in int32, vertical pass (sum + 2^21) >> 22, saturate to uint8.  The coefficient / offset tables
// are built on the host in float32 exactly like OpenCV (acr_b200/preprocess.py) and passed in, so the kernel
// is pure integer work and bit-reproducible.  (OpenCV builds that dispatch to IPP differ from this generic
// path by +-1 grey level on ~5 % of the pixels -- the reference itself is only defined up to that.)
#include "common.cuh"

namespace acr {

struct PreArgs {
  const uint8_t* src;   // (n, H, W, 3) BGR
  uint8_t* dst;         // (n, S, S, 3) RGB
  const int16_t* cx;    // (S,4) horizontal coefficients
  const int32_t* ox;    // (S) first-tap-plus-one source column in the PADDED square
  const int16_t* cy;
  const int32_t* oy;
  int n, H, W, S, side, pad_t, pad_l;
};

__global__ void __launch_bounds__(256) preprocess_kernel(PreArgs a) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)a.n * a.S * a.S;
  if (gid >= total) return;
  const int dx = (int)(gid % a.S), dy = (int)((gid / a.S) % a.S);
  const int img = (int)(gid / ((long long)a.S * a.S));
  const uint8_t* src = a.src + (size_t)img * a.H * a.W * 3;
  int acc[3] = {0, 0, 0};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int py = min(max(a.oy[dy] + j - 1, 0), a.side - 1);   // replicate border of the padded square
    const int sy = py - a.pad_t;
    int hor[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int px = min(max(a.ox[dx] + k - 1, 0), a.side - 1);
      const int sx = px - a.pad_l;
      const int w = a.cx[dx * 4 + k];
      if (sy >= 0 && sy < a.H && sx >= 0 && sx < a.W) {
        const uint8_t* p = src + ((size_t)sy * a.W + sx) * 3;
        hor[0] += w * p[2]; hor[1] += w * p[1]; hor[2] += w * p[0];   // BGR -> RGB
      } else {
        hor[0] += w * 255; hor[1] += w * 255; hor[2] += w * 255;      // white padding
      }
    }
    const int wy = a.cy[dy * 4 + j];
    acc[0] += wy * hor[0]; acc[1] += wy * hor[1]; acc[2] += wy * hor[2];
  }
  uint8_t* o = a.dst + (size_t)gid * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = (uint8_t)min(max((acc[c] + (1 << 21)) >> 22, 0), 255);
}

}  // namespace acr

using namespace acr;

extern "C" int acr_b200_preprocess(const uint8_t* frames_bgr, int n, int H, int W, const int16_t* coef_x,
                                   const int32_t* ofs_x, const int16_t* coef_y, const int32_t* ofs_y, int side,
                                   int pad_t, int pad_l, int out_size, uint8_t* out_rgb, void* stream) {
  ACR_CHECK_ARG(frames_bgr && out_rgb && coef_x && ofs_x && coef_y && ofs_y, "preprocess: null argument");
  ACR_CHECK_ARG(n > 0 && H > 0 && W > 0 && out_size > 0 && side >= H && side >= W && pad_t >= 0 && pad_l >= 0 &&
                    pad_t + H <= side && pad_l + W <= side, "preprocess: inconsistent geometry");
  PreArgs a{frames_bgr, out_rgb, coef_x, ofs_x, coef_y, ofs_y, n, H, W, out_size, side, pad_t, pad_l};
  const long long total = (long long)n * out_size * out_size;
  preprocess_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}
