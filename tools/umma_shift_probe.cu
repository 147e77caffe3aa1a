// Hardware probe: does a tcgen05.mma shared-memory descriptor accept a start address that is NOT aligned to the
// 1024-byte repeat of the 128-byte swizzle, and what does the descriptor's base-offset field (bits 49..51) do?
//
// One CTA loads ONE TMA box {64 ch, 24 px, 18 rows} (fp16, SWIZZLE_128B, 24-pixel row pitch = 3 swizzle repeats)
// and runs M=128 N=64 K=64 MMAs whose A descriptor starts `shift` pixels (128 B each) into the box, SBO = one
// image row (24 px), for every base offset 0..7.  B is the identity, so D[m][n] = A[row(m)][n]; the input encodes
// its own pixel index and 16-byte chunk id, so the output shows WHICH pixel and WHICH chunk order the tensor core
// read for every accumulator row.  Prints, per (shift, base offset), whether the rows are the expected
// (m/8)*24 + m%8 + shift pixels with chunks in order.  This decides whether the conv kernel can feed all nine taps
// of a 3x3 conv from a single haloed box (MODE_P1 in csrc/conv_tc.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_shift_probe tools/umma_shift_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 24)) { printf("probe: mbarrier timeout\n"); __trap(); }
  }
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

struct Params { CUtensorMap mA, mB; };
constexpr int PITCH = 24, ROWS = 18, NRUN_SHIFT = 6;
__constant__ int c_shift[NRUN_SHIFT] = {0, 1, 2, 8, 9, 25};   // pixels; 25 = one image row + 1 (tap ky=1,kx=1 of a 3x3)

// out[(si*8 + bo)][128][64] fp32
__global__ void __launch_bounds__(128) probe(const __grid_constant__ Params p, float* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base;                       // 18*24*128 = 55296 B
  const uint32_t b_smem = base + 55296;               // 64 rows x 128 B = 8192 B
  const uint32_t bar_ld = base + 55296 + 8192, bar_mma = bar_ld + 8, tptr = bar_ld + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_ld, 1); mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(raw + (tptr - smem_u32(raw)));
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_ld, 55296 + 8192);
    tma4(a_smem, &p.mA, bar_ld, 0, 0, 0, 0);
    tma2(b_smem, &p.mB, bar_ld, 0, 0);
  }
  mbar_wait(bar_ld, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // instruction descriptor: D=f32 (bit 4), A/B fp16 (0), K-major, N=64, M=128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t phase = 0;
  for (int si = 0; si < NRUN_SHIFT; ++si) {
    for (int bo = 0; bo < 8; ++bo) {
      if (threadIdx.x == 0) {
        const uint32_t a_start = a_smem + (uint32_t)c_shift[si] * 128u;
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t da = 0, db = 0;
          da |= (uint64_t)(((a_start + ks * 32) >> 4) & 0x3FFF);
          da |= (uint64_t)1 << 16;                                  // LBO (unused for swizzled K-major)
          da |= (uint64_t)(((PITCH * 128) >> 4) & 0x3FFF) << 32;    // SBO = one image row of the box
          da |= (uint64_t)1 << 46;                                  // version
          da |= (uint64_t)bo << 49;                                 // base offset under test
          da |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
          db |= (uint64_t)(((b_smem + ks * 32) >> 4) & 0x3FFF);
          db |= (uint64_t)1 << 16;
          db |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
          db |= (uint64_t)1 << 46;
          db |= (uint64_t)2 << 61;
          umma(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // warp w reads TMEM lanes 32w..32w+31 (accumulator rows), 64 columns
      float* o = out + ((size_t)(si * 8 + bo) * 128 + warp * 32 + lane) * 64;
      for (int c = 0; c < 64; c += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) o[c + i] = __uint_as_float(r[i]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  PFN enc = (PFN)fp;
  // A[y][x][c] fp16: c%8==0 -> pixel index y*24+x (exact in fp16 up to 2048), c%8==1 -> chunk id c/8, else 0
  std::vector<__half> hA((size_t)ROWS * PITCH * 64), hB(64 * 64);
  for (int y = 0; y < ROWS; ++y) for (int x = 0; x < PITCH; ++x) for (int c = 0; c < 64; ++c)
    hA[((size_t)y * PITCH + x) * 64 + c] = __float2half(c % 8 == 0 ? (float)(y * PITCH + x) : (c % 8 == 1 ? (float)(c / 8) : 0.f));
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2half(n == k ? 1.f : 0.f);
  __half *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2));
  const size_t no = (size_t)NRUN_SHIFT * 8 * 128 * 64;
  CK(cudaMalloc(&dO, no * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  Params p;
  cuuint32_t es[4] = {1, 1, 1, 1};
  {
    cuuint64_t dims[4] = {64, PITCH, ROWS, 1}, str[3] = {128, (cuuint64_t)PITCH * 128, (cuuint64_t)ROWS * PITCH * 128};
    cuuint32_t box[4] = {64, PITCH, ROWS, 1};
    CUresult r = enc(&p.mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {64, 64}, str[1] = {128};
    cuuint32_t box[2] = {64, 64};
    CUresult r = enc(&p.mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  const int smem = 55296 + 8192 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe<<<1, 128, smem>>>(p, dO);
  CK(cudaDeviceSynchronize());
  std::vector<float> hO(no);
  CK(cudaMemcpy(hO.data(), dO, no * 4, cudaMemcpyDeviceToHost));
  const int shifts[NRUN_SHIFT] = {0, 1, 2, 8, 9, 25};
  for (int si = 0; si < NRUN_SHIFT; ++si)
    for (int bo = 0; bo < 8; ++bo) {
      const float* o = hO.data() + (size_t)(si * 8 + bo) * 128 * 64;
      int bad_pix = 0, bad_chunk = 0, mixed = 0;
      for (int m = 0; m < 128; ++m) {
        const int want = (m / 8) * PITCH + (m % 8) + shifts[si];
        for (int j = 0; j < 8; ++j) {
          if ((int)o[m * 64 + 8 * j] != want) ++bad_pix;
          if ((int)o[m * 64 + 8 * j + 1] != j) ++bad_chunk;
          if ((int)o[m * 64 + 8 * j] != (int)o[m * 64]) ++mixed;
        }
      }
      printf("shift %2d base_offset %d : %s  (wrong pixel %4d/1024, wrong chunk order %4d/1024, rows mixing pixels %4d)", shifts[si], bo,
             (bad_pix == 0 && bad_chunk == 0) ? "OK   " : "WRONG", bad_pix, bad_chunk, mixed);
      // what did rows 0..9 read?  (pixel index per chunk position of row m, chunk ids of row m)
      if (bad_pix || bad_chunk) {
        printf("  rows0-2 pix:");
        for (int m = 0; m < 3; ++m) { printf(" ["); for (int j = 0; j < 8; ++j) printf("%d%s", (int)o[m * 64 + 8 * j], j < 7 ? "," : ""); printf("]"); }
        printf(" chunks row0: [");
        for (int j = 0; j < 8; ++j) printf("%d%s", (int)o[8 * j + 1], j < 7 ? "," : "");
        printf("] row7 pix: [");
        for (int j = 0; j < 8; ++j) printf("%d%s", (int)o[7 * 64 + 8 * j], j < 7 ? "," : "");
        printf("]");
      }
      printf("\n");
    }
  return 0;
}
