// Microbenchmark: TMA box-load throughput per SM for the box shapes the conv kernel uses.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench tools/tma_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

struct P { CUtensorMap m; int rank, stages, iters, box_bytes, W, H, N, tile_x, tile_y, rows2d; };

__global__ void __launch_bounds__(32) k(const __grid_constant__ P p, long long* out) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint32_t bars = base + p.stages * ((p.box_bytes + 1023) & ~1023);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long t0 = clock64();
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < p.iters; ++it) {
      if (it >= p.stages) mbar_wait(bars + 8 * s, ph ^ 1u);   // the load issued `stages` iterations ago
      mbar_expect_tx(bars + 8 * s, p.box_bytes);
      uint32_t dst = base + s * ((p.box_bytes + 1023) & ~1023);
      int t = blockIdx.x + it * gridDim.x;
      if (p.rank == 4) {
        int tx = t % (p.W / p.tile_x), ty = (t / (p.W / p.tile_x)) % (p.H / p.tile_y), n = (t / ((p.W / p.tile_x) * (p.H / p.tile_y))) % p.N;
        tma4(dst, &p.m, bars + 8 * s, 0, tx * p.tile_x - 1, ty * p.tile_y - 1, n);
      } else {
        tma2(dst, &p.m, bars + 8 * s, 0, (t * 128) % p.rows2d);
      }
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    // drain
    for (int d = 0; d < p.stages && d < p.iters; ++d) { mbar_wait(bars + 8 * s, ph ^ 1u); if (++s == p.stages) { s = 0; ph ^= 1u; } }
    out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  PFN enc = (PFN)fp;
  const int C = 64, W = 64, H = 64;
  long long* dout; CK(cudaMalloc(&dout, 148 * 8));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  for (int N : {4}) {
    size_t bytes = (size_t)N * H * W * C * 2;
    void* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 0, bytes));
    struct Cfg { const char* name; int rank; int bx, by; int cch; };
    Cfg cfgs[] = {{"4D {64,8,4,1}", 4, 8, 4, 64}, {"4D {64,8,8,1}", 4, 8, 8, 64}, {"4D {64,8,18,1}", 4, 8, 18, 64}, {"4D {64,8,32,1}", 4, 8, 32, 64},
                  {"4D {64,16,16,1}", 4, 16, 16, 64}, {"4D {32,16,18,1}", 4, 16, 18, 32}, {"4D {64,16,18,1}", 4, 16, 18, 64}, {"4D {64,32,18,1}", 4, 32, 18, 64}, {"4D {64,1,16,1}", 4, 1, 16, 64}};
    for (auto& c : cfgs) {
      for (int stages : {4}) {
        P p; p.rank = c.rank; p.stages = stages; p.iters = 400; p.W = W; p.H = H; p.N = N; p.rows2d = N * H * W;
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUtensorMapSwizzle sw = c.cch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        if (c.rank == 2) {
          cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)N * H * W}; cuuint64_t str[1] = {(cuuint64_t)C * 2}; cuuint32_t box[2] = {64, 128};
          if (enc(&p.m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc fail\n"); continue; }
          p.box_bytes = 64 * 128 * 2; p.tile_x = p.tile_y = 1;
        } else {
          cuuint64_t dims[4] = {(cuuint64_t)c.cch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
          cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
          cuuint32_t box[4] = {(cuuint32_t)c.cch, (cuuint32_t)c.bx, (cuuint32_t)c.by, 1};
          if (enc(&p.m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc fail %s\n", c.name); continue; }
          p.box_bytes = c.cch * c.bx * c.by * 2;
          p.tile_x = c.bx; p.tile_y = c.by >= 16 ? 16 : c.by;
        }
        size_t smem = stages * ((p.box_bytes + 1023) & ~1023) + 2048;
        k<<<148, 32, smem>>>(p, dout); CK(cudaDeviceSynchronize());
        k<<<148, 32, smem>>>(p, dout); CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        int rows = c.rank == 2 ? 128 : c.bx * c.by;
        printf("N=%3d %-26s stages %d: %8.0f clk/box  %6.1f clk/row  %6.2f B/clk/SM  (box %d B, %d rows)\n", N, c.name, stages, avg / p.iters, avg / p.iters / rows, p.box_bytes * (double)p.iters / avg, p.box_bytes, rows);
      }
    }
    CK(cudaFree(d));
  }
  return 0;
}
