// HBM-bound normalisation kernels on the frames-major channels-last bf16 layout.
//
//  * GroupNorm(32 groups) [+ SiLU]: replaces GroupNormSpecific / nn.GroupNorm + nn.SiLU at
//    openaimodel3d.py:151-155,175-180,256-266,555, attention.py:265,331, ae_modules.py:14-15,230-239.
//    Statistics are per (sample, group) over `rows_per_stat` consecutive rows: H*W rows for the
//    per-frame norms of ResBlock / SpatialTransformer, T*H*W rows for the 5-D norms of
//    TemporalConvBlock / TemporalTransformer (F.group_norm reduces over (C/G, T, H, W) there).
//    Two passes: (1) fp32 sum / sum-of-squares per group, one partial per block (no atomics, so the
//    result is deterministic); (2) y = x * scale[c] + shift[c] (+ SiLU) with scale/shift staged in
//    shared memory.  Algorithmic traffic: 2 reads + 1 write of the tensor (second read normally hits L2).
//  * LayerNorm over C per row: nn.LayerNorm at attention.py:229-231 (one warp per row, two-pass
//    in registers).
#include "common.cuh"
#include "geo4d_b200.h"

namespace g4 {

// ---------------------------------------------------------------------------------------------- GroupNorm
// block = (C/8, ty): each thread owns a fixed 8-channel vector and strides over rows.  Per-block partial
// {sum, sumsq} per group go to part[stat][block][32][2] (no atomics: deterministic); gn_apply sums the
// partials of its statistic while it builds the per-channel scale/shift table.
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int C, int rows_per_stat,
                                int rows_per_block, float* __restrict__ part /*[S][nblk][32][2]*/) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // [ty][2*C] partials, then [2*C] totals in row 0
  const int s = blockIdx.y;
  const int row0 = blockIdx.x * rows_per_block;
  const int row1 = min(row0 + rows_per_block, rows_per_stat);
  const int vec = threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sum[j] = 0.f; sq[j] = 0.f; }
  const __nv_bfloat16* base = x + ((long long)s * rows_per_stat) * ld + vec * 8;
  const int step = blockDim.y;
  int r = row0 + threadIdx.y;
  // 4 independent 16-byte loads in flight per thread
  for (; r + 3 * step < row1; r += 4 * step) {
    uint4 w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) w[u] = __ldg(reinterpret_cast<const uint4*>(base + (long long)(r + u * step) * ld));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 f;
      f = unpack_bf16x2(w[u].x); sum[0] += f.x; sq[0] += f.x * f.x; sum[1] += f.y; sq[1] += f.y * f.y;
      f = unpack_bf16x2(w[u].y); sum[2] += f.x; sq[2] += f.x * f.x; sum[3] += f.y; sq[3] += f.y * f.y;
      f = unpack_bf16x2(w[u].z); sum[4] += f.x; sq[4] += f.x * f.x; sum[5] += f.y; sq[5] += f.y * f.y;
      f = unpack_bf16x2(w[u].w); sum[6] += f.x; sq[6] += f.x * f.x; sum[7] += f.y; sq[7] += f.y * f.y;
    }
  }
  for (; r < row1; r += step) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(base + (long long)r * ld));
    float2 f;
    f = unpack_bf16x2(w.x); sum[0] += f.x; sq[0] += f.x * f.x; sum[1] += f.y; sq[1] += f.y * f.y;
    f = unpack_bf16x2(w.y); sum[2] += f.x; sq[2] += f.x * f.x; sum[3] += f.y; sq[3] += f.y * f.y;
    f = unpack_bf16x2(w.z); sum[4] += f.x; sq[4] += f.x * f.x; sum[5] += f.y; sq[5] += f.y * f.y;
    f = unpack_bf16x2(w.w); sum[6] += f.x; sq[6] += f.x * f.x; sum[7] += f.y; sq[7] += f.y * f.y;
  }
  float* mine = sm + (size_t)threadIdx.y * 2 * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[vec * 8 + j] = sum[j];
    mine[C + vec * 8 + j] = sq[j];
  }
  __syncthreads();
  for (int c = tid; c < 2 * C; c += nthr) {
    float a = 0.f;
    for (int y = 0; y < (int)blockDim.y; ++y) a += sm[(size_t)y * 2 * C + c];
    sm[c] = a;  // each column is read and written by exactly one thread
  }
  __syncthreads();
  const int cpg = C / 32;
  if (tid < 64) {
    const int g = tid & 31, which = tid >> 5;
    float a = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) a += sm[which * C + c];
    part[(((long long)s * gridDim.x + blockIdx.x) * 32 + g) * 2 + which] = a;
  }
}

__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ld, __nv_bfloat16* __restrict__ y,
                                long long ldy, int C, int rows_per_stat, int rows_per_block,
                                const float* __restrict__ part, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int silu) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // scale[C], shift[C], stats[64]
  const int s = blockIdx.y;
  const int row0 = blockIdx.x * rows_per_block;
  const int row1 = min(row0 + rows_per_block, rows_per_stat);
  const int vec = threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  const int cpg = C / 32;
  float* st = sm + 2 * C;
  if (tid < 64) {
    const int g = tid & 31, which = tid >> 5;
    float a = 0.f;
    for (int b = 0; b < (int)gridDim.x; ++b) a += part[(((long long)s * gridDim.x + b) * 32 + g) * 2 + which];
    st[which * 32 + g] = a;
  }
  __syncthreads();
  const float inv_cnt = 1.0f / ((float)cpg * (float)rows_per_stat);
  for (int c = tid; c < C; c += nthr) {
    const int g = c / cpg;
    const float mean = st[g] * inv_cnt;
    const float var = fmaxf(st[32 + g] * inv_cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float sc = gamma[c] * rstd;
    sm[c] = sc;
    sm[C + c] = beta[c] - mean * sc;
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = sm[vec * 8 + j]; sh[j] = sm[C + vec * 8 + j]; }
  const __nv_bfloat16* xb = x + ((long long)s * rows_per_stat) * ld + vec * 8;
  __nv_bfloat16* yb = y + ((long long)s * rows_per_stat) * ldy + vec * 8;
  const int step = blockDim.y;
  for (int r0 = row0 + threadIdx.y; r0 < row1; r0 += 4 * step) {
    uint4 w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r0 + u * step < row1) w[u] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)(r0 + u * step) * ld));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (r0 + u * step < row1) {
        float v[8];
        float2 f;
        f = unpack_bf16x2(w[u].x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(w[u].y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(w[u].z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(w[u].w); v[6] = f.x; v[7] = f.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = v[j] * sc[j] + sh[j];
          if (silu) v[j] = silu_f(v[j]);
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(yb + (long long)(r0 + u * step) * ldy) = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// one warp per row; C multiple of 8, C <= 8*32*MAXV
template <int MAXV>
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ld, __nv_bfloat16* __restrict__ y,
                                 long long ldy, int M, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps) {
  pdl_grid_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nvec = C >> 3;
  float v[MAXV][8];
  float sum = 0.f;
  const __nv_bfloat16* xr = x + (long long)warp * ld;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(xr + vi * 8));
      float2 f;
      f = unpack_bf16x2(w.x); v[i][0] = f.x; v[i][1] = f.y;
      f = unpack_bf16x2(w.y); v[i][2] = f.x; v[i][3] = f.y;
      f = unpack_bf16x2(w.z); v[i][4] = f.x; v[i][5] = f.y;
      f = unpack_bf16x2(w.w); v[i][6] = f.x; v[i][7] = f.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq += d * d; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  __nv_bfloat16* yr = y + (long long)warp * ldy;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      float o[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
      o[0] = (v[i][0] - mean) * rstd * g0.x + b0.x; o[1] = (v[i][1] - mean) * rstd * g0.y + b0.y;
      o[2] = (v[i][2] - mean) * rstd * g0.z + b0.z; o[3] = (v[i][3] - mean) * rstd * g0.w + b0.w;
      o[4] = (v[i][4] - mean) * rstd * g1.x + b1.x; o[5] = (v[i][5] - mean) * rstd * g1.y + b1.y;
      o[6] = (v[i][6] - mean) * rstd * g1.z + b1.z; o[7] = (v[i][7] - mean) * rstd * g1.w + b1.w;
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
      w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(yr + vi * 8) = w;
    }
  }
}

int device_sm_count();

}  // namespace g4

using namespace g4;

static void gn_launch_shape(int num_stats, int rows_per_stat, int C, int sms, dim3* block, dim3* grid,
                            int* rows_per_block) {
  const int vecs = C / 8;
  int ty = 512 / vecs; if (ty < 1) ty = 1; if (ty > 32) ty = 32;
  *block = dim3(vecs, ty);
  int blocks_per_stat = (4 * sms + num_stats - 1) / num_stats;  // ~4 blocks per SM overall
  int rpb = (rows_per_stat + blocks_per_stat - 1) / blocks_per_stat;
  if (rpb < 8 * ty) rpb = 8 * ty;
  blocks_per_stat = (rows_per_stat + rpb - 1) / rpb;
  *rows_per_block = rpb;
  *grid = dim3(blocks_per_stat, num_stats);
}

extern "C" size_t geo4d_groupnorm_workspace_bytes(int num_stats, int rows_per_stat, int C) {
  if (num_stats < 1 || rows_per_stat < 1 || C < 8) return 0;
  dim3 block, grid; int rpb;
  gn_launch_shape(num_stats, rows_per_stat, C, 148, &block, &grid, &rpb);  // upper bound: fewer SMs -> fewer blocks
  return (size_t)num_stats * grid.x * 64 * sizeof(float);
}

extern "C" int geo4d_groupnorm_silu(const void* x, int64_t ldx, void* y, int64_t ldy, int num_stats,
                                    int rows_per_stat, int C, const float* gamma, const float* beta, float eps,
                                    int apply_silu, void* workspace, size_t workspace_bytes, g4_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !y || !gamma || !beta || !workspace) { set_last_error("groupnorm: null pointer"); return G4_ERR_BAD_ARG; }
  if (C % 32 || C % 8 || C > 8 * 1024 || num_stats < 1 || num_stats > 65535 || rows_per_stat < 1) {
    set_last_error("groupnorm: C=%d must be a multiple of 32 and 8 (<=8192); num_stats=%d rows=%d", C, num_stats,
                   rows_per_stat);
    return G4_ERR_BAD_ARG;
  }
  if (ldx % 8 || ldy % 8 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) {
    set_last_error("groupnorm: x/y must be 16-byte aligned with ld multiple of 8"); return G4_ERR_BAD_ARG;
  }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  dim3 block, grid; int rows_per_block;
  gn_launch_shape(num_stats, rows_per_stat, C, sms > 148 ? 148 : sms, &block, &grid, &rows_per_block);
  const size_t need = (size_t)num_stats * grid.x * 64 * sizeof(float);
  if (workspace_bytes < need) { set_last_error("groupnorm: workspace %zu < %zu", workspace_bytes, need); return G4_ERR_WORKSPACE; }
  const size_t smem_stats = (size_t)block.y * 2 * C * sizeof(float);
  const size_t smem_apply = (2 * (size_t)C + 64) * sizeof(float);
  if (smem_stats > 48 * 1024) {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      if (e != cudaSuccess) { set_last_error("groupnorm: smem attr: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
      attr = true;
    }
  }
  launch_pdl(gn_stats_kernel, dim3(grid), dim3(block), smem_stats, stream, reinterpret_cast<const __nv_bfloat16*>(x), ldx, C, rows_per_stat,
                                                       rows_per_block, reinterpret_cast<float*>(workspace));
  int rc = check_launch("gn_stats"); if (rc) return rc;
  launch_pdl(gn_apply_kernel, dim3(grid), dim3(block), smem_apply, stream, reinterpret_cast<const __nv_bfloat16*>(x), ldx,
                                                       reinterpret_cast<__nv_bfloat16*>(y), ldy, C, rows_per_stat,
                                                       rows_per_block, reinterpret_cast<const float*>(workspace), gamma,
                                                       beta, eps, apply_silu);
  return check_launch("gn_apply");
}

extern "C" int geo4d_layernorm(const void* x, int64_t ldx, void* y, int64_t ldy, int M, int C, const float* gamma,
                               const float* beta, float eps, g4_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !y || !gamma || !beta) { set_last_error("layernorm: null pointer"); return G4_ERR_BAD_ARG; }
  if (C % 8 || C < 8 || C > 8 * 32 * 8 || M < 1) { set_last_error("layernorm: C=%d (multiple of 8, <=2048), M=%d", C, M); return G4_ERR_BAD_ARG; }
  if (ldx % 8 || ldy % 8 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(gamma) & 15) || (reinterpret_cast<uintptr_t>(beta) & 15)) {
    set_last_error("layernorm: pointers must be 16-byte aligned, ld multiple of 8"); return G4_ERR_BAD_ARG;
  }
  const int warps_per_block = 8;
  const int grid = (M + warps_per_block - 1) / warps_per_block;
  const int nvec = C / 8;
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  if (nvec <= 32) launch_pdl(layernorm_kernel<1>, dim3(grid), dim3(256), 0, stream, xp, ldx, yp, ldy, M, C, gamma, beta, eps);
  else if (nvec <= 64) launch_pdl(layernorm_kernel<2>, dim3(grid), dim3(256), 0, stream, xp, ldx, yp, ldy, M, C, gamma, beta, eps);
  else if (nvec <= 128) launch_pdl(layernorm_kernel<4>, dim3(grid), dim3(256), 0, stream, xp, ldx, yp, ldy, M, C, gamma, beta, eps);
  else launch_pdl(layernorm_kernel<8>, dim3(grid), dim3(256), 0, stream, xp, ldx, yp, ldy, M, C, gamma, beta, eps);
  return check_launch("layernorm");
}
