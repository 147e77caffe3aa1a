// Fused MANO forward for sm_100a: Rodrigues -> pose/shape blend shapes -> joint regression ->
// kinematic chain -> linear blend skinning -> fingertips / joint reorder / centring ->
// weak-perspective projection, one kernel, params in -> vertices / joints out.
//
// Replaces (reference, /root/reference): ManoLayer.forward mano/manolayer.py:104-276,
// batch_rodrigues :423-434, quat2mat :396-421, MANOWrapper.forward acr/mano_wrapper.py:37-50,
// batch_orth_proj / convert_kp2d_from_input_to_orgimg acr/utils.py:384-397.
//
// Work decomposition: CTA = (group of HG hands) x (chunk of VPB vertices).  Phase 1 (all 128
// threads, thread = (hand, joint)) rebuilds the 16 rigid transforms of each hand of the group in
// shared memory -- it is ~1% of the work, so every vertex chunk recomputes it instead of taking
// a second launch or a grid sync.  Phase 2 (thread = vertex) streams the 145 blend-shape rows
// once per CTA from L2 (coalesced, [k][c][v] layout) and applies them to all HG hands from
// registers, with the pose-map coefficients broadcast from shared memory as float4.
#include "common.cuh"
#include "rotation.cuh"

namespace acr {

constexpr int NV = 778;
constexpr int NVP = 784;   // vertex stride in the packed model (zero padded)
constexpr int NK = 145;    // 135 pose-map rows + 10 shape rows
constexpr int HG = 8;      // hands per CTA
constexpr int VPB = 128;   // vertices per CTA == threads per CTA

constexpr size_t OFF_DIRS = 0;                              // [NK][3][NVP]
constexpr size_t OFF_VT = OFF_DIRS + (size_t)NK * 3 * NVP;  // [3][NVP]
constexpr size_t OFF_W = OFF_VT + 3 * NVP;                  // [16][NVP]
constexpr size_t OFF_JT = OFF_W + 16 * NVP;                 // [16][3]   J_regressor . v_template
constexpr size_t OFF_JS = OFF_JT + 48;                      // [16][3][10] J_regressor . shapedirs
constexpr size_t OFF_HM = OFF_JS + 480;                     // [48] 0,0,0, hands_mean
constexpr size_t MODEL_FLOATS = OFF_HM + 48;

// out-joint index of source joint s (inverse of the reference's reorder list, manolayer.py:254)
__constant__ int c_joint_inv[21] = {0, 5, 6, 7, 9, 10, 11, 17, 18, 19, 13, 14, 15, 1, 2, 3, 4, 8, 12, 16, 20};
// source joint of out-joint i (the reference's list itself)
__constant__ int c_joint_perm[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};

struct ManoParams {
  const float* model[2];  // [0]=left, [1]=right
  const float* poses;
  const float* betas;
  const int32_t* hand_type;
  int default_side;
  const int32_t* n_dev;
  int n_max;
  int center_src;  // source-joint index (0..15) used as centre, -1 none
  const float* cam;
  const float* offsets;
  float* verts;
  float* joints;
  float* center;
  float* verts_camed;
  float* pj2d;
  float* pj2d_org;
  // fused vertex all-gather (acr_b200_mano_forward_gather, protocol in include/acr_b200.h)
  int gather;                    // 0 = plain launch
  char* peer_base[8];            // every rank's symmetric allocation, mapped into this rank's address space
  char* mc_base;                 // NVLS multicast address of the same allocation, or nullptr (-> per-peer stores)
  int world, rank;
  long long dst_row;             // first row of this rank's block inside a slot (= rank * rows, even)
  unsigned long long slot_bytes, counts_offset, flags_offset;
  unsigned long long* step_dev;  // device-local: gather launches completed so far
  unsigned int* done_ctr;        // device-local: CTAs of the running launch that have finished
  const int32_t* counts_src;     // (8) int32 row counts of this shard (acr_b200_parse), carried along
};

// one value / one 16-byte vector to every GPU of the multicast group: the NVSwitch replicates the store (NVLS)
__device__ __forceinline__ void multimem_st_f32(float* mc_addr, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_v4(float* mc_addr, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// bounded spin on arrival flags in THIS rank's memory: a peer that never arrives must fail the launch, not hang the box
__device__ __forceinline__ void wait_flag_ge(const unsigned long long* flag, unsigned long long target) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < target) {
    __nanosleep(200);
    if (clock64() - t0 > 60000000000ll) {   // ~30 s
      printf("acr_b200 gather: rank flag %p stuck at %llu < %llu\n", (const void*)flag, ld_acquire_sys(flag), target);
      __trap();
    }
  }
}

// packed fp32 pairs: Blackwell issues two IEEE fp32 FMAs per FFMA2 instruction (fma.rn.f32x2), which is what
// lifts the blend-shape loop off the FP32 issue bound (each lane is a plain fmaf: results are bit-identical)
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void unpk2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

constexpr int STAGE_ROW = 4 + VPB * 3;   // floats: a hand's 384 vertex floats at the offset that matches their global alignment

struct ProjCtx {  // per-hand projection constants kept in shared memory
  float s, tx, ty, padw, padh, ltx, lty;
};

__device__ __forceinline__ void write_joint(const ManoParams& p, int hand, int out_idx, float x, float y,
                                            float z, const ProjCtx& pc) {
  if (p.joints) {
    float* o = p.joints + ((size_t)hand * 21 + out_idx) * 3;
    o[0] = x; o[1] = y; o[2] = z;
  }
  if (p.cam) {
    float px = x * pc.s + pc.tx, py = y * pc.s + pc.ty;
    if (p.pj2d) {
      float* o = p.pj2d + ((size_t)hand * 21 + out_idx) * 2;
      o[0] = px; o[1] = py;
    }
    if (p.pj2d_org && p.offsets) {
      float* o = p.pj2d_org + ((size_t)hand * 21 + out_idx) * 2;
      o[0] = (px + 1.f) * pc.padw / 2.f + pc.ltx;
      o[1] = (py + 1.f) * pc.padh / 2.f + pc.lty;
    }
  }
}

__global__ void __launch_bounds__(VPB) mano_forward_kernel(const ManoParams p) {
  __shared__ __align__(16) float s_pm[NK][HG];        // pose-map / beta coefficients, [k][hand]
  __shared__ __align__(16) float s_A[HG][16][12];     // skinning transforms (rest pose removed)
  __shared__ float s_R[HG][16][9];                    // local rotations
  __shared__ float s_J[HG][16][3];                    // rest joints
  __shared__ float s_G[HG][16][12];                   // global transforms
  __shared__ float s_ctr[HG][3];
  __shared__ ProjCtx s_pc[HG];
  __shared__ int s_side[HG];
  __shared__ __align__(16) float s_stage[HG][STAGE_ROW];   // vertices of this CTA, staged for 16-byte stores
  __shared__ int s_last;

  const int n = p.n_dev ? min(*p.n_dev, p.n_max) : p.n_max;
  const int g0 = blockIdx.x * HG;
  const int t = threadIdx.x;
  // ---- fused all-gather bookkeeping.  This launch is gather step `step + 1`; it writes slot (step + 1) & 1.
  unsigned long long step = 0;
  char* slot_mc = nullptr;
  unsigned long long slot_off = 0;
  if (p.gather) {
    step = *p.step_dev;                       // stable during the launch: only the last CTA advances it, at the very end
    slot_off = ((step + 1) & 1ull) * p.slot_bytes;
    slot_mc = p.mc_base ? p.mc_base + slot_off : nullptr;
    // CTA (0,0) carries the 32 bytes of row counts of this shard to every rank (no separate collective)
    if (blockIdx.x == 0 && blockIdx.y == 0 && t < 8 && p.counts_src) {
      // the slot is free once every peer has signalled step `step` (see below)
      for (int r = 0; r < p.world; ++r) wait_flag_ge(reinterpret_cast<const unsigned long long*>(p.peer_base[p.rank] + p.flags_offset) + r, step);
      const int32_t cv = p.counts_src[t];
      const unsigned long long off = slot_off + p.counts_offset + ((size_t)p.rank * 8 + t) * 4;
      if (p.mc_base) multimem_st_f32(reinterpret_cast<float*>(p.mc_base + off), __int_as_float(cv));
      else for (int r = 0; r < p.world; ++r) *reinterpret_cast<int32_t*>(p.peer_base[r] + off) = cv;
    }
  }
  const bool active = g0 < n;
  if (active && p.gather) {
    // Slot (step+1)&1 was last written by step-1.  A peer signals step `step` only after its launch `step`, which it
    // enqueued after consuming step-1's data (the consume-before-next-launch contract) -> once every peer's flag in
    // OUR memory reads >= step, nobody reads the slot any more.  With two slots this wait is one whole step old.
    if (t < p.world) wait_flag_ge(reinterpret_cast<const unsigned long long*>(p.peer_base[p.rank] + p.flags_offset) + t, step);
    __syncthreads();
  }
  if (active) {

  // ---------------------------------------------------------------- phase 1: rigid transforms
  {
    const int h = t >> 4, j = t & 15;
    const int hand = g0 + h;
    const bool valid = hand < n;
    int side = p.default_side;
    if (valid && p.hand_type) side = p.hand_type[hand] != 0;
    if (!valid) side = -1;
    if (j == 0) {
      s_side[h] = side;
      ProjCtx pc = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid && p.cam) {
        pc.s = p.cam[hand * 3 + 0]; pc.tx = p.cam[hand * 3 + 1]; pc.ty = p.cam[hand * 3 + 2];
        if (p.offsets) {  // [pad_h,pad_w | crop t,r,b,l | pad t,r,b,l]  (acr/utils.py:392-397)
          const float* o = p.offsets + (size_t)hand * 10;
          pc.padw = o[0]; pc.padh = o[1];  // kp2d.x scales with offsets[:,0], kp2d.y with [:,1]
          pc.ltx = o[5] - o[9];
          pc.lty = o[2] - o[6];
        }
      }
      s_pc[h] = pc;
    }
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float J[3] = {0, 0, 0};
    if (valid) {
      const float* m = p.model[side];
      const float* ps = p.poses + (size_t)hand * 48 + j * 3;
      float ax = ps[0] + m[OFF_HM + j * 3 + 0];
      float ay = ps[1] + m[OFF_HM + j * 3 + 1];
      float az = ps[2] + m[OFF_HM + j * 3 + 2];
      rodrigues(ax, ay, az, R);
      const float* b = p.betas + (size_t)hand * 10;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float a = m[OFF_JT + j * 3 + c];
        const float* js = m + OFF_JS + (j * 3 + c) * 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) a = fmaf(js[k], b[k], a);
        J[c] = a;
      }
      if (j < 10) s_pm[135 + j][h] = b[j];
    } else if (j < 10) {
      s_pm[135 + j][h] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) s_R[h][j][e] = R[e];
#pragma unroll
    for (int c = 0; c < 3; ++c) s_J[h][j][c] = J[c];
    if (j >= 1) {
#pragma unroll
      for (int e = 0; e < 9; ++e) s_pm[(j - 1) * 9 + e][h] = valid ? R[e] - ((e & 3) == 0 ? 1.f : 0.f) : 0.f;
    }
    __syncthreads();
    // kinematic chain: thread j<5 walks finger j (joints 3j+1..3j+3); thread j==5 stores the root
    if (j <= 5) {
      float G[12];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        G[r * 4 + 0] = s_R[h][0][r * 3 + 0]; G[r * 4 + 1] = s_R[h][0][r * 3 + 1];
        G[r * 4 + 2] = s_R[h][0][r * 3 + 2]; G[r * 4 + 3] = s_J[h][0][r];
      }
      if (j == 5) {
#pragma unroll
        for (int e = 0; e < 12; ++e) s_G[h][0][e] = G[e];
      } else {
        int parent = 0;
#pragma unroll
        for (int lev = 0; lev < 3; ++lev) {
          const int idx = 3 * j + 1 + lev;
          const float* Rl = s_R[h][idx];
          float rel[3] = {s_J[h][idx][0] - s_J[h][parent][0], s_J[h][idx][1] - s_J[h][parent][1],
                          s_J[h][idx][2] - s_J[h][parent][2]};
          float N[12];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              N[r * 4 + c] = G[r * 4 + 0] * Rl[0 * 3 + c] + G[r * 4 + 1] * Rl[1 * 3 + c] + G[r * 4 + 2] * Rl[2 * 3 + c];
            N[r * 4 + 3] = G[r * 4 + 0] * rel[0] + G[r * 4 + 1] * rel[1] + G[r * 4 + 2] * rel[2] + G[r * 4 + 3];
          }
#pragma unroll
          for (int e = 0; e < 12; ++e) { G[e] = N[e]; s_G[h][idx][e] = N[e]; }
          parent = idx;
        }
      }
    }
    __syncthreads();
    // A_j = [R_g | t_g - R_g . J_j]   (manolayer.py:226-228)
    {
      const float* G = s_G[h][j];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        s_A[h][j][r * 4 + 0] = G[r * 4 + 0]; s_A[h][j][r * 4 + 1] = G[r * 4 + 1]; s_A[h][j][r * 4 + 2] = G[r * 4 + 2];
        s_A[h][j][r * 4 + 3] = G[r * 4 + 3] - (G[r * 4 + 0] * J[0] + G[r * 4 + 1] * J[1] + G[r * 4 + 2] * J[2]);
      }
    }
    if (j < 3) s_ctr[h][j] = (p.center_src >= 0) ? s_G[h][p.center_src][j * 4 + 3] : 0.f;
    __syncthreads();
    // kinematic joints + centre are written once per hand group (vertex chunk 0)
    if (blockIdx.y == 0 && valid) {
      const float* G = s_G[h][j];
      write_joint(p, hand, c_joint_inv[j], G[3] - s_ctr[h][0], G[7] - s_ctr[h][1], G[11] - s_ctr[h][2], s_pc[h]);
      if (j < 3 && p.center) p.center[(size_t)hand * 3 + j] = s_ctr[h][j];
    }
  }

  // ------------------------------------------------------------------ phase 2: vertices
  const int v = blockIdx.y * VPB + t;
  const bool vvalid = v < NV;
  const int vc = vvalid ? v : NVP - 1;  // padded (zero) column for the idle tail threads
  for (int side = 0; side < 2; ++side) {
    bool any = false;
#pragma unroll
    for (int h = 0; h < HG; ++h) any |= (s_side[h] == side);
    if (!any) continue;  // block-uniform
    const float* __restrict__ m = p.model[side];
    const float* __restrict__ dirs = m + OFF_DIRS;
    unsigned long long acc2[HG / 2][3];   // (hand 2i, hand 2i+1) packed: one FFMA2 serves two hands
    {
      const float v0 = m[OFF_VT + 0 * NVP + vc], v1 = m[OFF_VT + 1 * NVP + vc], v2 = m[OFF_VT + 2 * NVP + vc];
#pragma unroll
      for (int h = 0; h < HG / 2; ++h) { acc2[h][0] = pk2(v0, v0); acc2[h][1] = pk2(v1, v1); acc2[h][2] = pk2(v2, v2); }
    }
    // shape rows first (v_shaped), then pose rows, like the reference's evaluation order
#pragma unroll 5
    for (int kk = 0; kk < NK; ++kk) {
      const int k = (kk < 10) ? 135 + kk : kk - 10;
      const float d0 = __ldg(dirs + ((size_t)k * 3 + 0) * NVP + vc);
      const float d1 = __ldg(dirs + ((size_t)k * 3 + 1) * NVP + vc);
      const float d2 = __ldg(dirs + ((size_t)k * 3 + 2) * NVP + vc);
      const float4 pa = *reinterpret_cast<const float4*>(&s_pm[k][0]);
      const float4 pb = *reinterpret_cast<const float4*>(&s_pm[k][4]);
      const unsigned long long D0 = pk2(d0, d0), D1 = pk2(d1, d1), D2 = pk2(d2, d2);
      const unsigned long long P[4] = {pk2(pa.x, pa.y), pk2(pa.z, pa.w), pk2(pb.x, pb.y), pk2(pb.z, pb.w)};
#pragma unroll
      for (int h = 0; h < HG / 2; ++h) {
        ffma2(acc2[h][0], D0, P[h]);
        ffma2(acc2[h][1], D1, P[h]);
        ffma2(acc2[h][2], D2, P[h]);
      }
    }
    float acc[HG][3];
#pragma unroll
    for (int h = 0; h < HG / 2; ++h)
#pragma unroll
      for (int c = 0; c < 3; ++c) unpk2(acc2[h][c], acc[2 * h][c], acc[2 * h + 1][c]);
    float w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = __ldg(m + OFF_W + j * NVP + vc);
    // fingertip slot of this vertex (manolayer.py:244-247), -1 if none
    int tip = -1;
    if (v == 745) tip = 0; else if (v == 317) tip = 1; else if (v == (side ? 444 : 445)) tip = 2;
    else if (v == 556) tip = 3; else if (v == 673) tip = 4;
#pragma unroll
    for (int h = 0; h < HG; ++h) {
      if (s_side[h] != side) continue;  // block-uniform
      float T[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 a0 = *reinterpret_cast<const float4*>(&s_A[h][j][0]);
        const float4 a1 = *reinterpret_cast<const float4*>(&s_A[h][j][4]);
        const float4 a2 = *reinterpret_cast<const float4*>(&s_A[h][j][8]);
        T[0] = fmaf(w[j], a0.x, T[0]); T[1] = fmaf(w[j], a0.y, T[1]); T[2] = fmaf(w[j], a0.z, T[2]); T[3] = fmaf(w[j], a0.w, T[3]);
        T[4] = fmaf(w[j], a1.x, T[4]); T[5] = fmaf(w[j], a1.y, T[5]); T[6] = fmaf(w[j], a1.z, T[6]); T[7] = fmaf(w[j], a1.w, T[7]);
        T[8] = fmaf(w[j], a2.x, T[8]); T[9] = fmaf(w[j], a2.y, T[9]); T[10] = fmaf(w[j], a2.z, T[10]); T[11] = fmaf(w[j], a2.w, T[11]);
      }
      const float x = T[0] * acc[h][0] + T[1] * acc[h][1] + T[2] * acc[h][2] + T[3] - s_ctr[h][0];
      const float y = T[4] * acc[h][0] + T[5] * acc[h][1] + T[6] * acc[h][2] + T[7] - s_ctr[h][1];
      const float z = T[8] * acc[h][0] + T[9] * acc[h][1] + T[10] * acc[h][2] + T[11] - s_ctr[h][2];
      if (!vvalid) continue;
      const int hand = g0 + h;
      {   // stage at the offset that matches the global alignment of this hand's chunk (phase 3)
        const int o = (int)((((size_t)hand * NV + (size_t)blockIdx.y * VPB) * 3) & 3);
        float* sp = &s_stage[h][o + 3 * t];
        sp[0] = x; sp[1] = y; sp[2] = z;
      }
      if (tip >= 0) write_joint(p, hand, c_joint_inv[16 + tip], x, y, z, s_pc[h]);
    }
  }

  // ---------------------------------------------------------------- phase 3: 16-byte vertex stores
  // A hand's chunk is 3*nv contiguous floats at global float index gidx = (hand*778 + v0)*3, which is 8-byte but
  // not always 16-byte aligned (778*3 = 2 mod 4).  It was staged at offset (gidx & 3) of its shared-memory row, so
  // 16-byte chunk j of the row IS an aligned 16-byte chunk of global memory: one float4 (or multimem.st.v4 / peer
  // float4) per thread and chunk; the first / last chunk of a row may be partial and falls back to scalar stores.
  __syncthreads();
  {
    const int v0 = blockIdx.y * VPB;
    const int nfl = min(VPB, NV - v0) * 3;
    for (int h = 0; h < HG; ++h) {
      if (s_side[h] < 0) continue;                      // block-uniform
      const int hand = g0 + h;
      const size_t gidx = ((size_t)hand * NV + v0) * 3;
      const int o = (int)(gidx & 3);
      const size_t gg = ((size_t)(p.dst_row + hand) * NV + v0) * 3;      // same (gg & 3): dst_row is even
      const int nchunk = (o + nfl + 3) >> 2;
      const float sc = s_pc[h].s, tx = s_pc[h].tx, ty = s_pc[h].ty;
      for (int j = t; j < nchunk; j += VPB) {
        const float4 q = *reinterpret_cast<const float4*>(&s_stage[h][4 * j]);
        const float e[4] = {q.x, q.y, q.z, q.w};
        const int i0 = 4 * j - o;                       // float index inside the chunk of element 0 (may be < 0)
        const bool full = i0 >= 0 && i0 + 4 <= nfl;
        float c4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int comp = (i0 + k + 3) % 3;            // i0 + k >= -3
          c4[k] = comp == 0 ? e[k] * sc + tx : (comp == 1 ? e[k] * sc + ty : e[k]);
        }
        if (full) {
          if (p.verts) *reinterpret_cast<float4*>(p.verts + gidx + i0) = q;
          if (p.verts_camed && p.cam) *reinterpret_cast<float4*>(p.verts_camed + gidx + i0) = make_float4(c4[0], c4[1], c4[2], c4[3]);
          if (p.gather) {
            const unsigned long long off = slot_off + (gg + i0) * 4;
            if (slot_mc) multimem_st_v4(reinterpret_cast<float*>(p.mc_base + off), q);
            else for (int r = 0; r < p.world; ++r) *reinterpret_cast<float4*>(p.peer_base[r] + off) = q;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = i0 + k;
            if (i < 0 || i >= nfl) continue;
            if (p.verts) p.verts[gidx + i] = e[k];
            if (p.verts_camed && p.cam) p.verts_camed[gidx + i] = c4[k];
            if (p.gather) {
              const unsigned long long off = slot_off + (gg + i) * 4;
              if (slot_mc) multimem_st_f32(reinterpret_cast<float*>(p.mc_base + off), e[k]);
              else for (int r = 0; r < p.world; ++r) *reinterpret_cast<float*>(p.peer_base[r] + off) = e[k];
            }
          }
        }
      }
    }
  }
  }  // if (active)

  // ---------------------------------------------------------------- gather: completion signal
  // Every CTA (also the ones beyond n) counts itself done after a system-scope fence; the last one to finish
  // publishes step+1 in the flag word flags[rank] of EVERY rank (release, system scope): whoever acquires that
  // value sees all vertices and counts of this launch.  No separate barrier kernel, no NCCL call.
  if (p.gather) {
    __syncthreads();
    if (t == 0) {
      __threadfence_system();
      const unsigned int total = gridDim.x * gridDim.y;
      s_last = (atomicAdd(p.done_ctr, 1u) == total - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      if (t == 0) { __threadfence_system(); *p.done_ctr = 0u; }
      __syncthreads();
      if (t < p.world)
        st_release_sys(reinterpret_cast<unsigned long long*>(p.peer_base[t] + p.flags_offset) + p.rank, step + 1);
      if (t == 0) *p.step_dev = step + 1;
    }
  }
}

// stream-ordered wait until the stores of the most recent gather launch of EVERY rank have landed in this rank's memory
__global__ void gather_wait_kernel(const unsigned long long* flags, const unsigned long long* step_dev, int world) {
  if ((int)threadIdx.x < world) wait_flag_ge(flags + threadIdx.x, *step_dev);
}

// estimate_translation_np (acr/utils.py:430-472) for one hand per thread, fp64 like the numpy original:
// rows [f,0,cx-u | (u-cx)*Z - f*X] and [0,f,cy-v | (v-cy)*Z - f*Y] of every usable joint, normal equations.
__global__ void cam_trans_kernel(const float* __restrict__ j3d, const float* __restrict__ pj2d,
                                 const int32_t* __restrict__ n_dev, int n_max, float focal, float img_size,
                                 float* __restrict__ out) {
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double f = focal, c0 = (double)(img_size * 0.5f);
  double A00 = 0, A01 = 0, A02 = 0, A11 = 0, A12 = 0, A22 = 0, b0 = 0, b1 = 0, b2 = 0;
  int used = 0;
  for (int j = 0; j < 21; ++j) {
    const float X = j3d[((size_t)i * 21 + j) * 3 + 0], Y = j3d[((size_t)i * 21 + j) * 3 + 1], Z = j3d[((size_t)i * 21 + j) * 3 + 2];
    const float u = (pj2d[((size_t)i * 21 + j) * 2 + 0] + 1.f) * (img_size * 0.5f);
    const float v = (pj2d[((size_t)i * 21 + j) * 2 + 1] + 1.f) * (img_size * 0.5f);
    if (!(v > -2.f) || Z == -2.f) continue;   // the reference's "confidence" tests (acr/utils.py:489-492)
    ++used;
    const double qx = c0 - (double)u, qy = c0 - (double)v;        // third column of the two rows
    const double cx = ((double)u - c0) * (double)Z - f * (double)X, cy = ((double)v - c0) * (double)Z - f * (double)Y;
    A00 += f * f; A02 += f * qx; b0 += f * cx;
    A11 += f * f; A12 += f * qy; b1 += f * cy;
    A22 += qx * qx + qy * qy; b2 += qx * cx + qy * cy;
  }
  float* o = out + (size_t)i * 3;
  if (used < 4) { o[0] = o[1] = o[2] = -1.f; return; }
  // symmetric 3x3 solve (A01 = 0): Cramer's rule in fp64
  const double det = A00 * (A11 * A22 - A12 * A12) - A01 * (A01 * A22 - A12 * A02) + A02 * (A01 * A12 - A11 * A02);
  const double d0 = b0 * (A11 * A22 - A12 * A12) - A01 * (b1 * A22 - A12 * b2) + A02 * (b1 * A12 - A11 * b2);
  const double d1 = A00 * (b1 * A22 - A12 * b2) - b0 * (A01 * A22 - A12 * A02) + A02 * (A01 * b2 - b1 * A02);
  const double d2 = A00 * (A11 * b2 - b1 * A12) - A01 * (A01 * b2 - b1 * A02) + b0 * (A01 * A12 - A11 * A02);
  o[0] = (float)(d0 / det); o[1] = (float)(d1 / det); o[2] = (float)(d2 / det);
}

__global__ void rodrigues_kernel(const float* __restrict__ aa, int n, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float R[9];
  rodrigues(aa[i * 3 + 0], aa[i * 3 + 1], aa[i * 3 + 2], R);
#pragma unroll
  for (int e = 0; e < 9; ++e) out[(size_t)i * 9 + e] = R[e];
}

__global__ void rot6d_to_aa_kernel(const float* __restrict__ r6, int n, float* __restrict__ aa) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float o[3];
  rot6d_to_aa(r6 + (size_t)i * 6, o);
  aa[i * 3 + 0] = o[0]; aa[i * 3 + 1] = o[1]; aa[i * 3 + 2] = o[2];
}

}  // namespace acr

using namespace acr;

extern "C" size_t acr_b200_mano_model_floats(void) { return MODEL_FLOATS; }

extern "C" int acr_b200_mano_pack_model(const float* shapedirs, const float* posedirs, const float* v_template,
                                        const float* j_regressor, const float* weights,
                                        const float* hands_mean, int flip_x, float* out) {
  ACR_CHECK_ARG(shapedirs && posedirs && v_template && j_regressor && weights && hands_mean && out,
                "mano_pack_model: null argument");
  for (size_t i = 0; i < MODEL_FLOATS; ++i) out[i] = 0.f;
  auto sd = [&](int v, int c, int k) {
    float s = shapedirs[((size_t)v * 3 + c) * 10 + k];
    return (flip_x && c == 0) ? -s : s;
  };
  for (int v = 0; v < NV; ++v)
    for (int c = 0; c < 3; ++c) {
      for (int k = 0; k < 135; ++k)
        out[OFF_DIRS + ((size_t)k * 3 + c) * NVP + v] = posedirs[((size_t)v * 3 + c) * 135 + k];
      for (int k = 0; k < 10; ++k) out[OFF_DIRS + ((size_t)(135 + k) * 3 + c) * NVP + v] = sd(v, c, k);
      out[OFF_VT + (size_t)c * NVP + v] = v_template[v * 3 + c];
    }
  for (int v = 0; v < NV; ++v)
    for (int j = 0; j < 16; ++j) out[OFF_W + (size_t)j * NVP + v] = weights[v * 16 + j];
  // J = Jreg . (S.beta + T) = (Jreg.S).beta + Jreg.T   -- accumulate in double, store fp32
  for (int j = 0; j < 16; ++j)
    for (int c = 0; c < 3; ++c) {
      double t = 0;
      for (int v = 0; v < NV; ++v) t += (double)j_regressor[(size_t)j * NV + v] * v_template[v * 3 + c];
      out[OFF_JT + j * 3 + c] = (float)t;
      for (int k = 0; k < 10; ++k) {
        double s = 0;
        for (int v = 0; v < NV; ++v) s += (double)j_regressor[(size_t)j * NV + v] * sd(v, c, k);
        out[OFF_JS + (j * 3 + c) * 10 + k] = (float)s;
      }
    }
  for (int i = 0; i < 45; ++i) out[OFF_HM + 3 + i] = hands_mean[i];
  return ACR_B200_OK;
}

static int mano_forward_impl(const float* model_l, const float* model_r, const float* poses,
                             const float* betas, const int32_t* hand_type, int default_side,
                             const int32_t* n_dev, int n_max, int center_idx, const float* cam,
                             const float* offsets, float* verts, float* joints, float* center,
                             float* verts_camed, float* pj2d, float* pj2d_org, const int32_t* counts_src,
                             const acr_b200_gather* g, void* stream) {
  ACR_CHECK_ARG(n_max >= 0, "mano_forward: n_max < 0");
  if (n_max == 0 && !g) return ACR_B200_OK;
  ACR_CHECK_ARG(n_max > 0, "mano_forward_gather: every rank must launch every step (n_max > 0)");
  ACR_CHECK_ARG(poses && betas, "mano_forward: poses/betas are null");
  ACR_CHECK_ARG(default_side == 0 || default_side == 1, "mano_forward: default_side must be 0 or 1");
  ACR_CHECK_ARG(hand_type ? (model_l && model_r) : (default_side ? model_r != nullptr : model_l != nullptr),
                "mano_forward: missing packed model for a requested side");
  ACR_CHECK_ARG(center_idx >= -1 && center_idx < 21, "mano_forward: center_idx out of range");
  ACR_CHECK_ARG(((uintptr_t)verts | (uintptr_t)verts_camed) % 16 == 0, "mano_forward: verts / verts_camed must be 16-byte aligned");
  static const int perm[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};
  int center_src = -1;
  if (center_idx >= 0) {
    center_src = perm[center_idx];
    if (center_src >= 16) {
      set_error("mano_forward: centring on a fingertip joint (center_idx=%d) is not supported", center_idx);
      return ACR_B200_ENOTSUP;
    }
  }
  ManoParams p = {};
  p.model[0] = model_l ? model_l : model_r;
  p.model[1] = model_r ? model_r : model_l;
  p.poses = poses; p.betas = betas; p.hand_type = hand_type; p.default_side = default_side;
  p.n_dev = n_dev; p.n_max = n_max; p.center_src = center_src; p.cam = cam; p.offsets = offsets;
  p.verts = verts; p.joints = joints; p.center = center; p.verts_camed = verts_camed; p.pj2d = pj2d;
  p.pj2d_org = pj2d_org;
  if (g) {
    ACR_CHECK_ARG(g->world >= 1 && g->world <= 8 && g->rank >= 0 && g->rank < g->world, "mano_forward_gather: world / rank");
    ACR_CHECK_ARG(g->rows > 0 && g->rows % 2 == 0 && n_max <= g->rows, "mano_forward_gather: rows per rank must be even and >= n_max");
    ACR_CHECK_ARG(g->slot_bytes % 16 == 0 && g->counts_offset % 16 == 0 && g->flags_offset % 16 == 0 &&
                      g->counts_offset >= (uint64_t)g->world * g->rows * NV * 3 * 4 &&
                      g->slot_bytes >= g->counts_offset + (uint64_t)g->world * 32 && g->flags_offset >= 2 * g->slot_bytes,
                  "mano_forward_gather: slot layout");
    ACR_CHECK_ARG(g->local_state && (uintptr_t)g->local_state % 8 == 0, "mano_forward_gather: local_state");
    p.gather = 1; p.world = g->world; p.rank = g->rank;
    for (int r = 0; r < g->world; ++r) {
      ACR_CHECK_ARG(g->peer_base[r] && g->peer_base[r] % 16 == 0, "mano_forward_gather: peer base %d", r);
      p.peer_base[r] = reinterpret_cast<char*>(g->peer_base[r]);
    }
    p.mc_base = reinterpret_cast<char*>(g->multicast_base);
    p.dst_row = (long long)g->rank * g->rows;
    p.slot_bytes = g->slot_bytes; p.counts_offset = g->counts_offset; p.flags_offset = g->flags_offset;
    p.step_dev = reinterpret_cast<unsigned long long*>(g->local_state);
    p.done_ctr = reinterpret_cast<unsigned int*>(static_cast<char*>(g->local_state) + 8);
    p.counts_src = counts_src;
  }
  dim3 grid(ceil_div(n_max, HG), ceil_div(NV, VPB));
  mano_forward_kernel<<<grid, VPB, 0, (cudaStream_t)stream>>>(p);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

extern "C" int acr_b200_mano_forward(const float* model_l, const float* model_r, const float* poses,
                                     const float* betas, const int32_t* hand_type, int default_side,
                                     const int32_t* n_dev, int n_max, int center_idx, const float* cam,
                                     const float* offsets, float* verts, float* joints, float* center,
                                     float* verts_camed, float* pj2d, float* pj2d_org, void* stream) {
  return mano_forward_impl(model_l, model_r, poses, betas, hand_type, default_side, n_dev, n_max, center_idx, cam,
                           offsets, verts, joints, center, verts_camed, pj2d, pj2d_org, nullptr, nullptr, stream);
}

extern "C" int acr_b200_mano_forward_gather(const float* model_l, const float* model_r, const float* poses,
                                            const float* betas, const int32_t* hand_type, int default_side,
                                            const int32_t* n_dev, int n_max, int center_idx, const float* cam,
                                            const float* offsets, float* verts, float* joints, float* center,
                                            float* verts_camed, float* pj2d, float* pj2d_org,
                                            const int32_t* counts_src, const acr_b200_gather* gather, void* stream) {
  ACR_CHECK_ARG(gather != nullptr, "mano_forward_gather: gather descriptor is null");
  return mano_forward_impl(model_l, model_r, poses, betas, hand_type, default_side, n_dev, n_max, center_idx, cam,
                           offsets, verts, joints, center, verts_camed, pj2d, pj2d_org, counts_src, gather, stream);
}

extern "C" int acr_b200_gather_wait(const acr_b200_gather* g, void* stream) {
  ACR_CHECK_ARG(g && g->world >= 1 && g->world <= 8 && g->local_state && g->peer_base[g->rank], "gather_wait: bad descriptor");
  gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(g->peer_base[g->rank]) + g->flags_offset),
      reinterpret_cast<const unsigned long long*>(g->local_state), g->world);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

extern "C" int acr_b200_cam_trans(const float* j3d, const float* pj2d, const int32_t* n_dev, int n_max,
                                  float focal_length, float img_size, float* cam_trans, void* stream) {
  ACR_CHECK_ARG(n_max >= 0 && (n_max == 0 || (j3d && pj2d && cam_trans)), "cam_trans: bad arguments");
  if (n_max == 0) return ACR_B200_OK;
  cam_trans_kernel<<<ceil_div(n_max, 128), 128, 0, (cudaStream_t)stream>>>(j3d, pj2d, n_dev, n_max, focal_length, img_size, cam_trans);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

extern "C" int acr_b200_rot6d_to_aa(const float* rot6d, int n_rot, float* aa, void* stream) {
  ACR_CHECK_ARG(n_rot >= 0 && (n_rot == 0 || (rot6d && aa)), "rot6d_to_aa: bad arguments");
  if (n_rot == 0) return ACR_B200_OK;
  rot6d_to_aa_kernel<<<ceil_div(n_rot, 128), 128, 0, (cudaStream_t)stream>>>(rot6d, n_rot, aa);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}

extern "C" int acr_b200_rodrigues(const float* aa, int n_rot, float* rotmat, void* stream) {
  ACR_CHECK_ARG(n_rot >= 0 && (n_rot == 0 || (aa && rotmat)), "rodrigues: bad arguments");
  if (n_rot == 0) return ACR_B200_OK;
  rodrigues_kernel<<<ceil_div(n_rot, 128), 128, 0, (cudaStream_t)stream>>>(aa, n_rot, rotmat);
  ACR_CHECK_LAUNCH();
  return ACR_B200_OK;
}
