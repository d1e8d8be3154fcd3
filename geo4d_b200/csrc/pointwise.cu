// Small HBM-bound data-movement / element-wise kernels of the U-Net + sampler path.
#include "common.cuh"
#include "geo4d_b200.h"

namespace g4 {

// ---------------------------------------------------------------------------------------------- layout in
// fp32 5-D 'b c t h w' tensors (two sources concatenated along c, DiffusionWrapper 'hybrid'
// ddpm3d.py:2540-2544) -> bf16 rows [(b t h w), Cpad] zero padded.  One thread per pixel.
__global__ void bcthw_to_rows_kernel(const float* __restrict__ s0, int C0, const float* __restrict__ s1, int C1,
                                     int B, int T, int H, int W, __nv_bfloat16* __restrict__ out, int Cpad) {
  pdl_grid_sync();
  const long long npix = (long long)B * T * H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int x = i % W;
  const int y = (i / W) % H;
  const int t = (i / ((long long)W * H)) % T;
  const int b = i / ((long long)W * H * T);
  const long long thw = (long long)T * H * W;
  const long long off = ((long long)t * H + y) * W + x;
  __nv_bfloat16* o = out + i * Cpad;
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float f = 0.f;
      if (c < C0) f = s0[((long long)b * C0 + c) * thw + off];
      else if (c < C0 + C1) f = s1[((long long)b * C1 + (c - C0)) * thw + off];
      v[j] = f;
    }
    uint4 w;
    w.x = pack_bf16x2(v[0], v[1]); w.y = pack_bf16x2(v[2], v[3]);
    w.z = pack_bf16x2(v[4], v[5]); w.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(o + c0) = w;
  }
}

// fp32 rows [(b t h w), ld] (first C columns) -> fp32 'b c t h w'
__global__ void rows_to_bcthw_kernel(const float* __restrict__ rows, long long ld, int C, int B, int T, int H,
                                     int W, float* __restrict__ out) {
  pdl_grid_sync();
  const long long npix = (long long)B * T * H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const long long thw = (long long)T * H * W;
  const int b = i / thw;
  const long long off = i % thw;
  for (int c = 0; c < C; ++c) out[((long long)b * C + c) * thw + off] = rows[i * ld + c];
}

// ---------------------------------------------------------------------------------------------- concat / upsample / im2col
__global__ void concat_rows_kernel(const uint4* __restrict__ a, long long lda8, int va, const uint4* __restrict__ b,
                                   long long ldb8, int vb, uint4* __restrict__ out, long long rows) {
  pdl_grid_sync();
  const int vt = va + vb;
  const long long total = rows * vt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vt;
    const int v = i % vt;
    out[i] = (v < va) ? __ldg(a + r * lda8 + v) : __ldg(b + r * ldb8 + (v - va));
  }
}

__global__ void upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N, int H, int W,
                                  int vecs) {
  pdl_grid_sync();
  const long long total = (long long)N * (2 * H) * (2 * W) * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = i % vecs;
    const long long p = i / vecs;
    const int ox = p % (2 * W);
    const int oy = (p / (2 * W)) % (2 * H);
    const int n = p / ((long long)4 * W * H);
    out[i] = __ldg(in + (((long long)n * H + (oy >> 1)) * W + (ox >> 1)) * vecs + v);
  }
}

// 3x3 stride-2 im2col: out[(n, oy, ox), tap*C + c] = in[n, 2*oy + ky - pad, 2*ox + kx - pad, c] (0 outside)
__global__ void im2col_s2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N, int H, int W,
                                 int Ho, int Wo, int vecs, int pad) {
  pdl_grid_sync();
  const long long total = (long long)N * Ho * Wo * 9 * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = i % vecs;
    const int tap = (i / vecs) % 9;
    const long long p = i / ((long long)vecs * 9);
    const int ox = p % Wo;
    const int oy = (p / Wo) % Ho;
    const int n = p / ((long long)Wo * Ho);
    const int iy = 2 * oy + tap / 3 - pad;
    const int ix = 2 * ox + tap % 3 - pad;
    uint4 w = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) w = __ldg(in + (((long long)n * H + iy) * W + ix) * vecs + v);
    out[i] = w;
  }
}

// ---------------------------------------------------------------------------------------------- DDIM update
// p_sample_ddim, v-parameterisation, ddim.py:231-277 + ddpm3d.py:278-290, same fp32 operation order:
//   e_t   = sa * v + s1 * x
//   x0    = (sa * x - s1 * v) * rescale
//   x_prev= sqrt_a_prev * x0 + dir * e_t (+ sigma * noise)
// coef row = {sa, s1, rescale, sqrt_a_prev, dir, sigma}; the row index is read from *step_idx so the
// kernel can sit in a CUDA graph that is replayed once per step.
__global__ void ddim_step_kernel(float* __restrict__ x, const float* __restrict__ v, float* __restrict__ pred_x0,
                                 const float* __restrict__ noise, const float* __restrict__ coef,
                                 const int* __restrict__ step_idx, long long n) {
  pdl_grid_sync();
  const int s = step_idx ? *step_idx : 0;
  const float sa = coef[s * 6 + 0], s1 = coef[s * 6 + 1], rs = coef[s * 6 + 2];
  const float sap = coef[s * 6 + 3], dir = coef[s * 6 + 4], sg = coef[s * 6 + 5];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float xv = x[i], vv = v[i];
    const float e_t = __fadd_rn(__fmul_rn(sa, vv), __fmul_rn(s1, xv));
    float x0 = __fsub_rn(__fmul_rn(sa, xv), __fmul_rn(s1, vv));
    x0 = __fmul_rn(x0, rs);
    float xp = __fadd_rn(__fmul_rn(sap, x0), __fmul_rn(dir, e_t));
    if (noise) xp = __fadd_rn(xp, __fmul_rn(sg, noise[i]));
    x[i] = xp;
    if (pred_x0) pred_x0[i] = x0;
  }
}

__global__ void advance_counter_kernel(int* c, int delta, int modulo) {
  pdl_grid_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int v = *c + delta;
    if (modulo > 0) v %= modulo;
    *c = v;
  }
}

// out[j] = table[(*idx) * ld + j]  (per-step gather of precomputed embedding rows)
__global__ void gather_row_kernel(const float* __restrict__ table, long long ld, const int* __restrict__ idx,
                                  float* __restrict__ out, int n) {
  pdl_grid_sync();
  const int s = *idx;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    out[j] = table[(long long)s * ld + j];
}


// ---------------------------------------------------------------------------------------------- VAE AttnBlock helpers
// row softmax: fp32 scores -> bf16 probabilities (ae_modules.py:66-67); one warp per row.
__global__ void softmax_rows_kernel(const float* __restrict__ s, long long lds, __nv_bfloat16* __restrict__ p,
                                    long long ldp, long long rows, int cols) {
  pdl_grid_sync();
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* sr = s + row * lds;
  float mx = -INFINITY;
  for (int c = lane * 4; c < cols; c += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(sr + c));
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane * 4; c < cols; c += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(sr + c));
    sum += __expf(v.x - mx) + __expf(v.y - mx) + __expf(v.z - mx) + __expf(v.w - mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  __nv_bfloat16* pr = p + row * ldp;
  for (int c = lane * 4; c < cols; c += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(sr + c));
    uint2 w;
    w.x = pack_bf16x2(__expf(v.x - mx) * inv, __expf(v.y - mx) * inv);
    w.y = pack_bf16x2(__expf(v.z - mx) * inv, __expf(v.w - mx) * inv);
    *reinterpret_cast<uint2*>(pr + c) = w;
  }
}

// out[b, c, r] = in[b, r, c] (bf16), 32x32 tiles through shared memory
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ldin,
                                      __nv_bfloat16* __restrict__ out, int R, int Cc) {
  pdl_grid_sync();
  __shared__ __nv_bfloat16 tile[32][34];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const __nv_bfloat16* ib = in + (long long)b * R * ldin;
  __nv_bfloat16* ob = out + (long long)b * Cc * R;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = ib[(long long)r * ldin + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) ob[(long long)c * R + r] = tile[threadIdx.x][i];
  }
}

static inline int grid_for(long long total, int block, int cap) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int device_sm_count();

}  // namespace g4

using namespace g4;

#define G4_STREAM cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_)

extern "C" int geo4d_bcthw_to_rows(const float* src0, int C0, const float* src1, int C1, int B, int T, int H, int W,
                                   void* out, int Cpad, g4_stream_t stream_) {
  G4_STREAM;
  if (!src0 || !out || Cpad % 8 || Cpad < C0 + C1 || (C1 > 0 && !src1)) {
    set_last_error("bcthw_to_rows: bad args (Cpad=%d must be a multiple of 8 and >= C0+C1=%d)", Cpad, C0 + C1);
    return G4_ERR_BAD_ARG;
  }
  const long long npix = (long long)B * T * H * W;
  launch_pdl(bcthw_to_rows_kernel, dim3((int)((npix + 127) / 128)), dim3(128), 0, stream, src0, C0, src1, C1, B, T, H, W,
                                                                     reinterpret_cast<__nv_bfloat16*>(out), Cpad);
  return check_launch("bcthw_to_rows");
}

extern "C" int geo4d_rows_to_bcthw(const float* rows, int64_t ld, int C, int B, int T, int H, int W, float* out,
                                   g4_stream_t stream_) {
  G4_STREAM;
  if (!rows || !out) { set_last_error("rows_to_bcthw: null"); return G4_ERR_BAD_ARG; }
  const long long npix = (long long)B * T * H * W;
  launch_pdl(rows_to_bcthw_kernel, dim3((int)((npix + 127) / 128)), dim3(128), 0, stream, rows, ld, C, B, T, H, W, out);
  return check_launch("rows_to_bcthw");
}

extern "C" int geo4d_concat_rows(const void* a, int64_t lda, int Ca, const void* b, int64_t ldb, int Cb, void* out,
                                 int64_t rows, g4_stream_t stream_) {
  G4_STREAM;
  if (!a || !b || !out || Ca % 8 || Cb % 8 || lda % 8 || ldb % 8) { set_last_error("concat_rows: channels/ld must be multiples of 8"); return G4_ERR_BAD_ARG; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  const long long total = rows * ((Ca + Cb) / 8);
  launch_pdl(concat_rows_kernel, dim3(grid_for(total, 256, sms * 16)), dim3(256), 0, stream, 
      reinterpret_cast<const uint4*>(a), lda / 8, Ca / 8, reinterpret_cast<const uint4*>(b), ldb / 8, Cb / 8,
      reinterpret_cast<uint4*>(out), rows);
  return check_launch("concat_rows");
}

extern "C" int geo4d_upsample_nearest2x(const void* in, void* out, int N, int H, int W, int C, g4_stream_t stream_) {
  G4_STREAM;
  if (!in || !out || C % 8) { set_last_error("upsample2x: C must be a multiple of 8"); return G4_ERR_BAD_ARG; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  const long long total = (long long)N * 4 * H * W * (C / 8);
  launch_pdl(upsample2x_kernel, dim3(grid_for(total, 256, sms * 16)), dim3(256), 0, stream, 
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), N, H, W, C / 8);
  return check_launch("upsample2x");
}

extern "C" int geo4d_im2col_3x3_s2(const void* in, void* out, int N, int H, int W, int C, int pad_before,
                                   int Ho, int Wo, g4_stream_t stream_) {
  G4_STREAM;
  if (!in || !out || C % 8) { set_last_error("im2col: C must be a multiple of 8"); return G4_ERR_BAD_ARG; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  const long long total = (long long)N * Ho * Wo * 9 * (C / 8);
  launch_pdl(im2col_s2_kernel, dim3(grid_for(total, 256, sms * 16)), dim3(256), 0, stream, 
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), N, H, W, Ho, Wo, C / 8, pad_before);
  return check_launch("im2col_3x3_s2");
}

extern "C" int geo4d_ddim_step(float* x, const float* v, float* pred_x0, const float* noise, const float* coef,
                               const int* step_idx, int64_t n, g4_stream_t stream_) {
  G4_STREAM;
  if (!x || !v || !coef) { set_last_error("ddim_step: null"); return G4_ERR_BAD_ARG; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  launch_pdl(ddim_step_kernel, dim3(grid_for(n, 256, sms * 8)), dim3(256), 0, stream, x, v, pred_x0, noise, coef, step_idx, n);
  return check_launch("ddim_step");
}

extern "C" int geo4d_advance_counter(int* counter, int delta, int modulo, g4_stream_t stream_) {
  G4_STREAM;
  if (!counter) { set_last_error("advance_counter: null"); return G4_ERR_BAD_ARG; }
  launch_pdl(advance_counter_kernel, dim3(1), dim3(32), 0, stream, counter, delta, modulo);
  return check_launch("advance_counter");
}

extern "C" int geo4d_gather_row(const float* table, int64_t ld, const int* idx, float* out, int n,
                                g4_stream_t stream_) {
  G4_STREAM;
  if (!table || !idx || !out) { set_last_error("gather_row: null"); return G4_ERR_BAD_ARG; }
  launch_pdl(gather_row_kernel, dim3(grid_for(n, 256, 64)), dim3(256), 0, stream, table, ld, idx, out, n);
  return check_launch("gather_row");
}

extern "C" int geo4d_softmax_rows(const float* s, int64_t lds, void* p, int64_t ldp, int64_t rows, int cols,
                                  g4_stream_t stream_) {
  G4_STREAM;
  if (!s || !p || cols % 4 || lds % 4 || ldp % 4 || ((uintptr_t)s & 15) || ((uintptr_t)p & 7)) {
    set_last_error("softmax_rows: cols/ld must be multiples of 4 and pointers aligned"); return G4_ERR_BAD_ARG;
  }
  const long long blocks = (rows + 7) / 8;
  launch_pdl(softmax_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, s, lds, reinterpret_cast<__nv_bfloat16*>(p), ldp, rows, cols);
  return check_launch("softmax_rows");
}

extern "C" int geo4d_transpose_bf16(const void* in, int64_t ldin, void* out, int batch, int R, int Cc,
                                    g4_stream_t stream_) {
  G4_STREAM;
  if (!in || !out || batch < 1 || batch > 65535) { set_last_error("transpose: bad args"); return G4_ERR_BAD_ARG; }
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, batch), block(32, 8);
  launch_pdl(transpose_bf16_kernel, dim3(grid), dim3(block), 0, stream, reinterpret_cast<const __nv_bfloat16*>(in), ldin,
                                                    reinterpret_cast<__nv_bfloat16*>(out), R, Cc);
  return check_launch("transpose_bf16");
}
