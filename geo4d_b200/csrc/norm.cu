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
// {sum, sumsq} per group go to part[stat][block][32][2]; the LAST block of a statistic to finish (ticket)
// folds the partials in block order -- fixed order, no float atomics, so the result is deterministic --
// into fin[stat][64] = {mean[32], rstd[32]} and puts the ticket back to zero.  gn_apply reads 64 floats.
//
// workspace layout (floats): [0, GN_TICKETS) tickets (int, zero between calls) | fin[S][64] | part[S][nblk][64]
constexpr int GN_TICKETS = 4096;

__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int C, int rows_per_stat,
                                int rows_per_block, float* __restrict__ ws, float eps) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // [ty][2*C] partials, then [2*C] totals in row 0
  __shared__ int s_last;
  const int s = blockIdx.y;
  const int S = gridDim.y, nblk = gridDim.x;
  int* tickets = reinterpret_cast<int*>(ws);
  float* fin = ws + GN_TICKETS;
  float* part = fin + (size_t)S * 64;
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
    part[(((long long)s * nblk + blockIdx.x) * 32 + g) * 2 + which] = a;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&tickets[s], 1) == nblk - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // fold the partials of this statistic: thread (slice, pair) sums blocks slice, slice+nsl, ...; then the
  // slices are added in order.  Both orders are fixed by the launch shape.
  const int nsl = nthr >> 6;   // >= 1: the block has at least 64 threads
  if (tid < nsl * 64) {
    const int pair = tid & 63, slice = tid >> 6;
    const int g = pair & 31, which = pair >> 5;
    float a = 0.f;
    for (int b2 = slice; b2 < nblk; b2 += nsl) a += __ldcg(part + (((long long)s * nblk + b2) * 32 + g) * 2 + which);
    sm[slice * 64 + pair] = a;
  }
  __syncthreads();
  if (tid < 32) {
    float su = 0.f, sq2 = 0.f;
    for (int k = 0; k < nsl; ++k) { su += sm[k * 64 + tid]; sq2 += sm[k * 64 + 32 + tid]; }
    const float inv_cnt = 1.0f / ((float)cpg * (float)rows_per_stat);
    const float mean = su * inv_cnt;
    const float var = fmaxf(sq2 * inv_cnt - mean * mean, 0.f);
    fin[(size_t)s * 64 + tid] = mean;
    fin[(size_t)s * 64 + 32 + tid] = rsqrtf(var + eps);
  }
  if (tid == 0) tickets[s] = 0;
}

__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ld, __nv_bfloat16* __restrict__ y,
                                long long ldy, int C, int rows_per_stat, int rows_per_block,
                                const float* __restrict__ ws, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int silu) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // scale[C], shift[C]
  const int s = blockIdx.y;
  const float* fin = ws + GN_TICKETS + (size_t)s * 64;
  const int row0 = blockIdx.x * rows_per_block;
  const int row1 = min(row0 + rows_per_block, rows_per_stat);
  const int vec = threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  const int cpg = C / 32;
  for (int c = tid; c < C; c += nthr) {
    const int g = c / cpg;
    const float sc = gamma[c] * fin[32 + g];
    sm[c] = sc;
    sm[C + c] = beta[c] - fin[g] * sc;
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

// Single-launch variant: statistics, a per-statistic grid barrier, then the apply pass (which re-reads x,
// normally still in L2).  Launched cooperatively so that all blocks are co-resident; the barrier is
// "last block to arrive folds + publishes, everybody else polls one flag word".  Counters live in the first
// 16 KiB of the workspace (tickets | flags | leave counters | unused) and are back to zero when the kernel ends.
__global__ void gn_fused_kernel(const __nv_bfloat16* __restrict__ x, long long ld, __nv_bfloat16* __restrict__ y,
                                long long ldy, int C, int rows_per_stat, int rows_per_block, float* __restrict__ ws,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // [ty][2*C] partials / fold scratch, later scale[C] shift[C]
  __shared__ int s_last;
  const int s = blockIdx.y;
  const int S = gridDim.y, nblk = gridDim.x;
  int* tickets = reinterpret_cast<int*>(ws);
  int* flags = tickets + GN_TICKETS / 4;
  int* leaves = tickets + GN_TICKETS / 2;
  float* fin = ws + GN_TICKETS;
  float* part = fin + (size_t)S * 64;
  const int row0 = blockIdx.x * rows_per_block;
  const int row1 = min(row0 + rows_per_block, rows_per_stat);
  const int vec = threadIdx.x;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  const int cpg = C / 32;
  const int step = blockDim.y;
  const __nv_bfloat16* xb = x + ((long long)s * rows_per_stat) * ld + vec * 8;
  {
    float sum[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sum[j] = 0.f; sq[j] = 0.f; }
    for (int r0 = row0 + threadIdx.y; r0 < row1; r0 += 4 * step) {
      uint4 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        w[u] = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + u * step < row1) w[u] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)(r0 + u * step) * ld));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float2 f;
        f = unpack_bf16x2(w[u].x); sum[0] += f.x; sq[0] += f.x * f.x; sum[1] += f.y; sq[1] += f.y * f.y;
        f = unpack_bf16x2(w[u].y); sum[2] += f.x; sq[2] += f.x * f.x; sum[3] += f.y; sq[3] += f.y * f.y;
        f = unpack_bf16x2(w[u].z); sum[4] += f.x; sq[4] += f.x * f.x; sum[5] += f.y; sq[5] += f.y * f.y;
        f = unpack_bf16x2(w[u].w); sum[6] += f.x; sq[6] += f.x * f.x; sum[7] += f.y; sq[7] += f.y * f.y;
      }
    }
    float* mine = sm + (size_t)threadIdx.y * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[vec * 8 + j] = sum[j];
      mine[C + vec * 8 + j] = sq[j];
    }
  }
  __syncthreads();
  for (int c = tid; c < 2 * C; c += nthr) {
    float a = 0.f;
    for (int yy = 0; yy < (int)blockDim.y; ++yy) a += sm[(size_t)yy * 2 * C + c];
    sm[c] = a;
  }
  __syncthreads();
  if (tid < 64) {
    const int g = tid & 31, which = tid >> 5;
    float a = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) a += sm[which * C + c];
    part[(((long long)s * nblk + blockIdx.x) * 32 + g) * 2 + which] = a;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&tickets[s], 1) == nblk - 1) ? 1 : 0;
  __syncthreads();
  if (s_last) {
    __threadfence();
    const int nsl = nthr >> 6;
    if (tid < nsl * 64) {
      const int pair = tid & 63, slice = tid >> 6;
      const int g = pair & 31, which = pair >> 5;
      float a = 0.f;
      for (int b2 = slice; b2 < nblk; b2 += nsl) a += __ldcg(part + (((long long)s * nblk + b2) * 32 + g) * 2 + which);
      sm[slice * 64 + pair] = a;
    }
    __syncthreads();
    if (tid < 32) {
      float su = 0.f, sq2 = 0.f;
      for (int k = 0; k < nsl; ++k) { su += sm[k * 64 + tid]; sq2 += sm[k * 64 + 32 + tid]; }
      const float inv_cnt = 1.0f / ((float)cpg * (float)rows_per_stat);
      const float mean = su * inv_cnt;
      const float var = fmaxf(sq2 * inv_cnt - mean * mean, 0.f);
      fin[(size_t)s * 64 + tid] = mean;
      fin[(size_t)s * 64 + 32 + tid] = rsqrtf(var + eps);
      __threadfence();
    }
    __syncthreads();
    if (tid == 0) atomicExch(&flags[s], 1);
  } else if (tid == 0) {
    long long t0 = clock64();
    while (*reinterpret_cast<volatile int*>(&flags[s]) == 0) {
      __nanosleep(40);
      if (clock64() - t0 > G4_MBAR_TIMEOUT_CYCLES) {
        printf("g4: groupnorm grid barrier timeout (block %d,%d)\n", (int)blockIdx.x, (int)blockIdx.y);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
  // ---- apply
  for (int c = tid; c < C; c += nthr) {
    const int g = c / cpg;
    const float sc = gamma[c] * __ldcg(fin + (size_t)s * 64 + 32 + g);
    sm[c] = sc;
    sm[C + c] = beta[c] - __ldcg(fin + (size_t)s * 64 + g) * sc;
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = sm[vec * 8 + j]; sh[j] = sm[C + vec * 8 + j]; }
  __nv_bfloat16* yb = y + ((long long)s * rows_per_stat) * ldy + vec * 8;
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
  // ---- leave: the last block of this statistic puts the counters back to zero
  __syncthreads();
  if (tid == 0 && atomicAdd(&leaves[s], 1) == nblk - 1) {
    tickets[s] = 0;
    flags[s] = 0;
    leaves[s] = 0;
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// one warp per ROWS consecutive rows (all their loads are issued before any arithmetic, so a warp keeps
// ROWS * MAXV 16-byte requests in flight); C multiple of 8, C <= 8*32*MAXV; two-pass statistics in registers
template <int MAXV, int ROWS>
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ld, __nv_bfloat16* __restrict__ y,
                                 long long ldy, int M, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps) {
  pdl_grid_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = warp * ROWS;
  if (row0 >= M) return;
  const int nvec = C >> 3;
  uint4 w[ROWS][MAXV];
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    const __nv_bfloat16* xr = x + (long long)(row0 + rr) * ld;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      w[rr][i] = make_uint4(0u, 0u, 0u, 0u);
      if (vi < nvec && row0 + rr < M) w[rr][i] = __ldg(reinterpret_cast<const uint4*>(xr + vi * 8));
    }
  }
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    if (row0 + rr >= M) break;
    float v[MAXV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      float2 f;
      f = unpack_bf16x2(w[rr][i].x); v[i][0] = f.x; v[i][1] = f.y;
      f = unpack_bf16x2(w[rr][i].y); v[i][2] = f.x; v[i][3] = f.y;
      f = unpack_bf16x2(w[rr][i].z); v[i][4] = f.x; v[i][5] = f.y;
      f = unpack_bf16x2(w[rr][i].w); v[i][6] = f.x; v[i][7] = f.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];   // lanes beyond nvec hold zeros
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq += d * d; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
    __nv_bfloat16* yr = y + (long long)(row0 + rr) * ldy;
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
        uint4 q;
        q.x = pack_bf16x2(o[0], o[1]); q.y = pack_bf16x2(o[2], o[3]);
        q.z = pack_bf16x2(o[4], o[5]); q.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(yr + vi * 8) = q;
      }
    }
  }
}

int device_sm_count();
static bool g_gn_two_kernels = false;   // debug switch: keep statistics and apply in separate launches

}  // namespace g4

using namespace g4;

static void gn_launch_shape(int num_stats, int rows_per_stat, int C, int sms, dim3* block, dim3* grid,
                            int* rows_per_block) {
  const int vecs = C / 8;
  int ty = 512 / vecs; if (ty < 1) ty = 1; if (ty > 32) ty = 32;
  *block = dim3(vecs, ty);
  int blocks_per_stat = (3 * sms + num_stats - 1) / num_stats;  // ~3 blocks per SM overall
  int rpb = (rows_per_stat + blocks_per_stat - 1) / blocks_per_stat;
  if (rpb < 8 * ty) rpb = 8 * ty;
  blocks_per_stat = (rows_per_stat + rpb - 1) / rpb;
  *rows_per_block = rpb;
  *grid = dim3(blocks_per_stat, num_stats);
}

extern "C" void geo4d_debug_groupnorm_two_kernels(int on) { g4::g_gn_two_kernels = on != 0; }

extern "C" size_t geo4d_groupnorm_workspace_bytes(int num_stats, int rows_per_stat, int C) {
  if (num_stats < 1 || rows_per_stat < 1 || C < 8) return 0;
  dim3 block, grid; int rpb;
  gn_launch_shape(num_stats, rows_per_stat, C, 148, &block, &grid, &rpb);  // upper bound: fewer SMs -> fewer blocks
  return ((size_t)GN_TICKETS + (size_t)num_stats * 64 + (size_t)num_stats * grid.x * 64) * sizeof(float);
}

extern "C" int geo4d_groupnorm_silu(const void* x, int64_t ldx, void* y, int64_t ldy, int num_stats,
                                    int rows_per_stat, int C, const float* gamma, const float* beta, float eps,
                                    int apply_silu, void* workspace, size_t workspace_bytes, g4_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !y || !gamma || !beta || !workspace) { set_last_error("groupnorm: null pointer"); return G4_ERR_BAD_ARG; }
  if (C % 32 || C % 8 || C > 8 * 1024 || num_stats < 1 || num_stats > GN_TICKETS || rows_per_stat < 1) {
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
  size_t smem_stats = (size_t)block.y * 2 * C * sizeof(float);
  const size_t fold = (size_t)(block.x * block.y / 64) * 64 * sizeof(float);
  if (smem_stats < fold) smem_stats = fold;
  const size_t smem_apply = 2 * (size_t)C * sizeof(float);
  static unsigned long long attr_mask = 0;
  if (once_per_device(&attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) { unlatch_device(&attr_mask); set_last_error("groupnorm: smem attr: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  }
  // single cooperative launch when every block can be resident at once (always true for the U-Net shapes)
  if (!g_gn_two_kernels && num_stats <= GN_TICKETS / 4) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel, (int)(block.x * block.y), smem_stats);
    if (e == cudaSuccess && per_sm > 0) {
      const long long cap = (long long)per_sm * sms;
      dim3 g2 = grid;
      int rpb2 = rows_per_block;
      if ((long long)g2.x * g2.y > cap && cap >= num_stats) {
        g2.x = (unsigned)(cap / num_stats);
        rpb2 = (rows_per_stat + (int)g2.x - 1) / (int)g2.x;
        g2.x = (unsigned)((rows_per_stat + rpb2 - 1) / rpb2);
      }
      if ((long long)g2.x * g2.y <= cap) {
        const size_t need2 = ((size_t)GN_TICKETS + (size_t)num_stats * 64 + (size_t)num_stats * g2.x * 64) * sizeof(float);
        if (workspace_bytes < need2) { set_last_error("groupnorm: workspace %zu < %zu", workspace_bytes, need2); return G4_ERR_WORKSPACE; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = g2; cfg.blockDim = block; cfg.dynamicSmemBytes = smem_stats; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative;
        at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, gn_fused_kernel, reinterpret_cast<const __nv_bfloat16*>(x), (long long)ldx,
                               reinterpret_cast<__nv_bfloat16*>(y), (long long)ldy, C, rows_per_stat, rpb2,
                               reinterpret_cast<float*>(workspace), gamma, beta, eps, apply_silu);
        if (e == cudaSuccess) return check_launch("gn_fused");
        (void)cudaGetLastError();   // fall through to the two-kernel path
      }
    }
  }
  const size_t need = ((size_t)GN_TICKETS + (size_t)num_stats * 64 + (size_t)num_stats * grid.x * 64) * sizeof(float);
  if (workspace_bytes < need) { set_last_error("groupnorm: workspace %zu < %zu", workspace_bytes, need); return G4_ERR_WORKSPACE; }
  launch_pdl(gn_stats_kernel, dim3(grid), dim3(block), smem_stats, stream, reinterpret_cast<const __nv_bfloat16*>(x), ldx, C, rows_per_stat,
                                                       rows_per_block, reinterpret_cast<float*>(workspace), eps);
  int rc = check_launch("gn_stats"); if (rc) return rc;
  launch_pdl(gn_apply_kernel, dim3(grid), dim3(block), smem_apply, stream, reinterpret_cast<const __nv_bfloat16*>(x), ldx,
                                                       reinterpret_cast<__nv_bfloat16*>(y), ldy, C, rows_per_stat,
                                                       rows_per_block, reinterpret_cast<const float*>(workspace), gamma,
                                                       beta, apply_silu);
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
  const int nvec = C / 8;
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  auto grid_for_rows = [&](int rows_per_warp) { return (M + 8 * rows_per_warp - 1) / (8 * rows_per_warp); };
  if (nvec <= 32) launch_pdl(layernorm_kernel<1, 4>, dim3(grid_for_rows(4)), dim3(256), 0, stream, xp, ldx, yp, ldy, M, C, gamma, beta, eps);
  else if (nvec <= 64) launch_pdl(layernorm_kernel<2, 4>, dim3(grid_for_rows(4)), dim3(256), 0, stream, xp, ldx, yp, ldy, M, C, gamma, beta, eps);
  else if (nvec <= 128) launch_pdl(layernorm_kernel<4, 2>, dim3(grid_for_rows(2)), dim3(256), 0, stream, xp, ldx, yp, ldy, M, C, gamma, beta, eps);
  else launch_pdl(layernorm_kernel<8, 1>, dim3(grid_for_rows(1)), dim3(256), 0, stream, xp, ldx, yp, ldy, M, C, gamma, beta, eps);
  return check_launch("layernorm");
}
