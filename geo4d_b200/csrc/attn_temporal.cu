// Temporal self-attention over the frame axis (T <= 16 tokens per sequence, head dim 64).
//
// Replaces CrossAttention.forward (einsum + softmax path, attention.py:81-144) as used by
// TemporalTransformer (attention.py:365-412): for every pixel p and head h the 16 frame tokens
// attend to each other.  The reference materialises (b*hw*heads, 16, 16) score tensors and several
// rearrange copies; here one warp owns one (pixel, head) problem, stages its 3 x [T x 64] bf16 tiles
// in shared memory and runs  S = Q K^T  and  O = P V  on mma.sync.m16n8k16 fragments (a 16x16x64
// problem is far below one tcgen05 tile, and the op is HBM-bound: 4 x M x inner x 2 bytes).
// q/k/v are read straight out of the fused qkv projection (row stride ld, no head-split copies) and the
// output is written back in token-major [rows, inner] layout.
#include "common.cuh"
#include "geo4d_b200.h"

namespace g4 {

constexpr int TA_STRIDE = 72;  // 64 + 8 bf16 padding: 144-byte rows, ldmatrix conflict-free

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// rows are indexed row(b, t, p) = (b*T + t)*HW + p
__global__ void __launch_bounds__(128)
temporal_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ v, long long ld, __nv_bfloat16* __restrict__ o,
                     long long ldo, int B, int T, int HW, int heads, float scale_log2) {
  pdl_grid_sync();
  __shared__ __align__(16) __nv_bfloat16 sm[4][3][16 * TA_STRIDE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* sQ = sm[warp][0];
  __nv_bfloat16* sK = sm[warp][1];
  __nv_bfloat16* sV = sm[warp][2];
  const long long nprob = (long long)B * HW * heads;
  const int g = lane >> 2, qd = lane & 3;

  for (long long pid = (long long)blockIdx.x * 4 + warp; pid < nprob; pid += (long long)gridDim.x * 4) {
    const int h = pid % heads;
    const long long bp = pid / heads;
    const int p = bp % HW;
    const int b = bp / HW;
    // ---- stage Q, K, V (T rows x 8 16-byte vectors each); rows >= T are zero
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;
      const int t = idx >> 3, vec = idx & 7;
      uint4 wq = make_uint4(0, 0, 0, 0), wk = wq, wv = wq;
      if (t < T) {
        const long long off = ((long long)(b * T + t) * HW + p) * ld + h * 64 + vec * 8;
        wq = __ldg(reinterpret_cast<const uint4*>(q + off));
        wk = __ldg(reinterpret_cast<const uint4*>(k + off));
        wv = __ldg(reinterpret_cast<const uint4*>(v + off));
      }
      *reinterpret_cast<uint4*>(sQ + t * TA_STRIDE + vec * 8) = wq;
      *reinterpret_cast<uint4*>(sK + t * TA_STRIDE + vec * 8) = wk;
      *reinterpret_cast<uint4*>(sV + t * TA_STRIDE + vec * 8) = wv;
    }
    __syncwarp();
    // ---- S = Q K^T : two 16x8 score tiles, four k-steps of 16
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
      ldsm_x4(smem_u32(sQ + ((lane & 7) + 8 * ((lane >> 3) & 1)) * TA_STRIDE + kk * 16 + 8 * (lane >> 4)), a0, a1,
              a2, a3);
      ldsm_x4(smem_u32(sK + ((lane & 7) + 8 * (lane >> 4)) * TA_STRIDE + kk * 16 + 8 * ((lane >> 3) & 1)), b0, b1,
              b2, b3);
      mma_bf16_16816(s0, a0, a1, a2, a3, b0, b1);
      mma_bf16_16816(s1, a0, a1, a2, a3, b2, b3);
    }
    // ---- softmax over keys (each row lives in one quad): thread holds keys {2qd,2qd+1} (+8)
    float m_lo, m_hi, l_lo, l_hi;
    {
      const int j0 = 2 * qd, j1 = 2 * qd + 1, j2 = 8 + 2 * qd, j3 = 9 + 2 * qd;
      const float ninf = -INFINITY;
      float x0 = j0 < T ? s0[0] * scale_log2 : ninf, x1 = j1 < T ? s0[1] * scale_log2 : ninf;
      float x2 = j2 < T ? s1[0] * scale_log2 : ninf, x3 = j3 < T ? s1[1] * scale_log2 : ninf;
      float y0 = j0 < T ? s0[2] * scale_log2 : ninf, y1 = j1 < T ? s0[3] * scale_log2 : ninf;
      float y2 = j2 < T ? s1[2] * scale_log2 : ninf, y3 = j3 < T ? s1[3] * scale_log2 : ninf;
      m_lo = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
      m_hi = fmaxf(fmaxf(y0, y1), fmaxf(y2, y3));
      m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
      m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
      m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
      m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
      s0[0] = ex2_ftz(x0 - m_lo); s0[1] = ex2_ftz(x1 - m_lo); s1[0] = ex2_ftz(x2 - m_lo); s1[1] = ex2_ftz(x3 - m_lo);
      s0[2] = ex2_ftz(y0 - m_hi); s0[3] = ex2_ftz(y1 - m_hi); s1[2] = ex2_ftz(y2 - m_hi); s1[3] = ex2_ftz(y3 - m_hi);
      l_lo = s0[0] + s0[1] + s1[0] + s1[1];
      l_hi = s0[2] + s0[3] + s1[2] + s1[3];
      l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
      l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
      l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
      l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    }
    const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
    // P as the A fragment of the second GEMM (normalised here; fp32 -> bf16 like the tensor-core path)
    const uint32_t pa0 = pack_bf16x2(s0[0] * inv_lo, s0[1] * inv_lo);
    const uint32_t pa1 = pack_bf16x2(s0[2] * inv_hi, s0[3] * inv_hi);
    const uint32_t pa2 = pack_bf16x2(s1[0] * inv_lo, s1[1] * inv_lo);
    const uint32_t pa3 = pack_bf16x2(s1[2] * inv_hi, s1[3] * inv_hi);
    // ---- O = P V : eight 16x8 output tiles, one k-step (16 keys)
    __syncwarp();  // everyone is done reading sQ -> reuse it as the output staging tile
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(smem_u32(sV + ((lane & 7) + 8 * ((lane >> 3) & 1)) * TA_STRIDE + nn * 16 + 8 * (lane >> 4)), b0, b1,
                b2, b3);
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(o0, pa0, pa1, pa2, pa3, b0, b1);
      mma_bf16_16816(o1, pa0, pa1, pa2, pa3, b2, b3);
      *reinterpret_cast<uint32_t*>(sQ + g * TA_STRIDE + nn * 16 + 2 * qd) = pack_bf16x2(o0[0], o0[1]);
      *reinterpret_cast<uint32_t*>(sQ + (g + 8) * TA_STRIDE + nn * 16 + 2 * qd) = pack_bf16x2(o0[2], o0[3]);
      *reinterpret_cast<uint32_t*>(sQ + g * TA_STRIDE + nn * 16 + 8 + 2 * qd) = pack_bf16x2(o1[0], o1[1]);
      *reinterpret_cast<uint32_t*>(sQ + (g + 8) * TA_STRIDE + nn * 16 + 8 + 2 * qd) = pack_bf16x2(o1[2], o1[3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;
      const int t = idx >> 3, vec = idx & 7;
      if (t < T) {
        const long long off = ((long long)(b * T + t) * HW + p) * ldo + h * 64 + vec * 8;
        *reinterpret_cast<uint4*>(o + off) = *reinterpret_cast<const uint4*>(sQ + t * TA_STRIDE + vec * 8);
      }
    }
    __syncwarp();
  }
}

int device_sm_count();

}  // namespace g4

using namespace g4;

extern "C" int geo4d_temporal_attention(const void* q, const void* k, const void* v, int64_t ld, void* out,
                                        int64_t ldo, int B, int T, int HW, int heads, float scale,
                                        g4_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!q || !k || !v || !out) { set_last_error("temporal_attention: null pointer"); return G4_ERR_BAD_ARG; }
  if (T < 1 || T > 16) { set_last_error("temporal_attention: T=%d unsupported (1..16)", T); return G4_ERR_UNSUPPORTED; }
  if (ld % 8 || ldo % 8 || ((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15) || ((uintptr_t)out & 15)) {
    set_last_error("temporal_attention: pointers must be 16-byte aligned, ld multiple of 8"); return G4_ERR_BAD_ARG;
  }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  const long long nprob = (long long)B * HW * heads;
  long long grid = (nprob + 3) / 4;
  if (grid > (long long)sms * 16) grid = (long long)sms * 16;
  launch_pdl(temporal_attn_kernel, dim3((int)grid), dim3(128), 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), ld, reinterpret_cast<__nv_bfloat16*>(out), ldo, B, T, HW, heads,
      scale * 1.4426950408889634f);
  return check_launch("temporal_attention");
}
