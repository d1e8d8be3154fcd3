// Geometry kernels of the per-window post-processing and the sliding-window alignment
// (scripts/evaluation/infer_geo4d.py:447-500, utils/rays.py, dust3r/cloud_opt/*).  All fp32, HBM-bound:
// every kernel is a single coalesced pass over the point maps with warp-shuffle + block reductions.
#include "common.cuh"
#include "geo4d_b200.h"

namespace g4 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduce NV per-thread partial sums over the block and atomically add them to dst[0..NV) (fp64 accumulators).
template <int NV>
__device__ __forceinline__ void block_reduce_atomic(float (&acc)[NV], double* dst) {
  __shared__ float red[32][NV + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += (double)red[w][i];
    atomicAdd(dst + i, s);
    __threadfence();  // callers may elect a "last block" that reads the totals
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------- K8
// Per-window post-processing of the decoded maps (infer_geo4d.py:447-487): channels of `maps`
// [11][T][H][W] = pts xyz, raw confidence, ray dir xyz, ray moment xyz, inverse depth (mean of 3 already).
//   conf = softplus(c); invalid = sky(|v - 1.05| < eps on all 3) | far(any |v| > far_value)
//   inv_conf = invalid ? 0 : 1/conf; pts = (x/alpha, y/beta, (z+1)/2); invdepth = (d+1)/2
__global__ void postprocess_window_kernel(const float* __restrict__ maps, long long thw, float* __restrict__ pts,
                                          float* __restrict__ inv_conf, float* __restrict__ invd,
                                          unsigned char* __restrict__ valid, float sky_value, float sky_eps,
                                          float far_value, float alpha, float beta, int has_conf) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < thw; i += (long long)gridDim.x * blockDim.x) {
    const float x = maps[i], y = maps[thw + i], z = maps[2 * thw + i];
    float conf = 1.0f;
    if (has_conf) {
      const float c = maps[3 * thw + i];
      conf = (c > 20.0f) ? c : log1pf(__expf(c));  // nn.Softplus(beta=1, threshold=20)
    }
    const float lo = sky_value - sky_eps, hi = sky_value + sky_eps;
    const bool sky = (x > lo) && (x < hi) && (y > lo) && (y < hi) && (z > lo) && (z < hi);
    const bool far = (fabsf(x) > far_value) || (fabsf(y) > far_value) || (fabsf(z) > far_value);
    const bool invalid = sky || far;
    inv_conf[i] = invalid ? 0.0f : 1.0f / conf;
    pts[3 * i + 0] = x / alpha;
    pts[3 * i + 1] = y / beta;
    pts[3 * i + 2] = (z + 1.0f) / 2.0f;
    invd[i] = (maps[10 * thw + i] + 1.0f) / 2.0f;
    if (valid) valid[i] = invalid ? 0 : valid[i];
  }
}

// ---------------------------------------------------------------------------------------------- a14
// Ray map -> per-frame camera moments (utils/rays.py:301-367,387-433,579-595; utils/normalize.py:25-51).
// For frame t over the centre-cropped square: d = normalize(raydir), p = d x m,
//   M += I - d d^T (6 unique), b += (I - d d^T) p (3), Hm += d_t (x) d_0 (9)   -> out[t][18] (fp64)
__global__ void raymap_moments_kernel(const float* __restrict__ raydir, const float* __restrict__ raymom, int T,
                                      int H, int W, int x0, int y0, int S, double* __restrict__ out) {
  const int t = blockIdx.y;
  const long long hw = (long long)H * W, thw = hw * T;
  float acc[18];
#pragma unroll
  for (int i = 0; i < 18; ++i) acc[i] = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * S; i += gridDim.x * blockDim.x) {
    const int yy = y0 + i / S, xx = x0 + i % S;
    const long long o = (long long)t * hw + (long long)yy * W + xx;  // channel-major [3][T][H][W]
    const long long o0 = (long long)yy * W + xx;
    float dx = raydir[o], dy = raydir[thw + o], dz = raydir[2 * thw + o];
    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    dx *= inv; dy *= inv; dz *= inv;
    float ax = raydir[o0], ay = raydir[thw + o0], az = raydir[2 * thw + o0];
    const float ia = 1.0f / sqrtf(ax * ax + ay * ay + az * az);
    ax *= ia; ay *= ia; az *= ia;
    const float mx = raymom[o], my = raymom[thw + o], mz = raymom[2 * thw + o];
    const float px = dy * mz - dz * my, py = dz * mx - dx * mz, pz = dx * my - dy * mx;  // p = d x m
    // intersect_skew_lines_high_dim re-normalises r (already unit) -> I - d d^T
    const float c00 = 1.f - dx * dx, c01 = -dx * dy, c02 = -dx * dz, c11 = 1.f - dy * dy, c12 = -dy * dz,
                c22 = 1.f - dz * dz;
    acc[0] += c00; acc[1] += c01; acc[2] += c02; acc[3] += c11; acc[4] += c12; acc[5] += c22;
    acc[6] += c00 * px + c01 * py + c02 * pz;
    acc[7] += c01 * px + c11 * py + c12 * pz;
    acc[8] += c02 * px + c12 * py + c22 * pz;
    // H = B^T A with B = frame-t dirs, A = frame-0 dirs: H[i][j] = sum b_i a_j
    acc[9] += dx * ax;  acc[10] += dx * ay; acc[11] += dx * az;
    acc[12] += dy * ax; acc[13] += dy * ay; acc[14] += dy * az;
    acc[15] += dz * ax; acc[16] += dz * ay; acc[17] += dz * az;
  }
  block_reduce_atomic<18>(acc, out + (long long)t * 18);
}

// ---------------------------------------------------------------------------------------------- K9
// Affine map of point sets: set i (pts_per_set points) is mapped with its own row-major 3x4 matrix [A | t].
// mode 0: out[p] = A x[p] + t (3 floats); mode 1: out[p] = third row only (the camera-frame depth of
// init_from_pts3d_group, init_im_poses.py:612-620).  Replaces the small cuBLAS matmul / einsum of the initialisation.
__global__ void transform_points_kernel(const float* __restrict__ x, long long pts_per_set, const float* __restrict__ mats,
                                        float* __restrict__ out, int mode) {
  const int set = blockIdx.y;
  float M[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) M[i] = mats[set * 12 + i];
  const float* xs = x + (long long)set * pts_per_set * 3;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < pts_per_set; p += (long long)gridDim.x * blockDim.x) {
    const float a = xs[3 * p], b = xs[3 * p + 1], c = xs[3 * p + 2];
    if (mode == 0) {
      float* o = out + ((long long)set * pts_per_set + p) * 3;
      o[0] = M[0] * a + M[1] * b + M[2] * c + M[3];
      o[1] = M[4] * a + M[5] * b + M[6] * c + M[7];
      o[2] = M[8] * a + M[9] * b + M[10] * c + M[11];
    } else {
      out[(long long)set * pts_per_set + p] = M[8] * a + M[9] * b + M[10] * c + M[11];
    }
  }
}

// Weighted Umeyama moments (what roma.rigid_points_registration reduces; init_im_poses.py:797-800).
// pass 0: out[0..7)  = {sum w, sum w x(3), sum w y(3)}
// pass 1: out[0..10) = {sum w |x-xm|^2, sum w (y-ym)(x-xm)^T (9, row-major [i=y][j=x])}  given means[6]
__global__ void umeyama_moments_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                       const float* __restrict__ w1, const float* __restrict__ w2, long long n,
                                       int pass, const double* __restrict__ means, double* __restrict__ out) {
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  float xm0 = 0, xm1 = 0, xm2 = 0, ym0 = 0, ym1 = 0, ym2 = 0;
  if (pass == 1) {
    xm0 = (float)means[0]; xm1 = (float)means[1]; xm2 = (float)means[2];
    ym0 = (float)means[3]; ym1 = (float)means[4]; ym2 = (float)means[5];
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w = w1[i];
    if (w2) w *= w2[i];
    const float a0 = x[3 * i], a1 = x[3 * i + 1], a2 = x[3 * i + 2];
    const float b0 = y[3 * i], b1 = y[3 * i + 1], b2 = y[3 * i + 2];
    if (pass == 0) {
      acc[0] += w;
      acc[1] += w * a0; acc[2] += w * a1; acc[3] += w * a2;
      acc[4] += w * b0; acc[5] += w * b1; acc[6] += w * b2;
    } else {
      const float p0 = a0 - xm0, p1 = a1 - xm1, p2 = a2 - xm2;
      const float q0 = b0 - ym0, q1 = b1 - ym1, q2 = b2 - ym2;
      acc[0] += w * (p0 * p0 + p1 * p1 + p2 * p2);
      acc[1] += w * q0 * p0; acc[2] += w * q0 * p1; acc[3] += w * q0 * p2;
      acc[4] += w * q1 * p0; acc[5] += w * q1 * p1; acc[6] += w * q1 * p2;
      acc[7] += w * q2 * p0; acc[8] += w * q2 * p1; acc[9] += w * q2 * p2;
    }
  }
  block_reduce_atomic<10>(acc, out);
}

// ---------------------------------------------------------------------------------------------- K10
// One fused iteration of the global alignment objective (optimizer_group.py:440-525) for the dense
// part: per image n and pixel p
//   X_w = R_n (d (u-cx)/f, d (v-cy)/f, d) + T_n,  d = exp(logd[n][p])
//   L1 += w ||X_w - S_g pred_e[p]|| / A  over the edges e = (g, frame) that observe image n
//   L2 += 2 m |1/(d+1e-6) - (s_g rho_e[p] + t_g)| / A                    (phase B, it >= 150)
// It applies the Adam update to logd[n][p] in place (torch.optim.Adam formula, betas (0.9, 0.9)) and
// reduces the gradients w.r.t. the small parameters in matrix form:
//   gpose[n][12] = dL/d[R_n | T_n],  gS[g][12] = dL/d(s_g [R_g | T_g]),  gscal = {dL/d(1/f), L1, L2},
//   gst[g][2] = {dL/ds_g, dL/dt_g}
// The chain rule to quaternions / log-scales and the O(N) pose terms run on the host-side module.
constexpr int AL_KMAX = 8;

struct AlignIterArgs {
  float* logd; float* adam_m; float* adam_v;  // [N][HW]
  const float* pred;    // [E][HW][3]
  const float* weight;  // [E][HW]
  const float* invd;    // [E][HW] or null
  const int* edge_ptr;  // [N+1]
  const int* edge_idx;  // edges incident to each image
  const float* poses;   // [N][12] c2w rows
  const float* S;       // [G][12]
  const float* scal;    // [iters][8] rows {1/f slot (unused), cx, cy, lr, bias_corr1, bias_corr2_sqrt, inv_area, phaseB}
  const float* invf;    // [1] current 1/f
  const int* it;        // device iteration counter selecting the scal row (null -> row 0)
  const float* st;      // [G][3] {s, t, depth_valid}
  double* gpose; double* gS; double* gscal; double* gst;
  int N, HW, W, group_size;
};

__global__ void __launch_bounds__(256)
align_iter_kernel(const AlignIterArgs a) {
  const int n = blockIdx.y;
  const int e0 = a.edge_ptr[n], ne = min(a.edge_ptr[n + 1] - e0, AL_KMAX);
  const float* sc = a.scal + (a.it ? (long long)(*a.it) * 8 : 0);
  const float invf = a.invf[0], cx = sc[1], cy = sc[2], lr = sc[3], bc1 = sc[4], bc2s = sc[5], invA = sc[6];
  const bool phaseB = sc[7] != 0.f && a.invd != nullptr;
  float R[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) R[i] = a.poses[n * 12 + i];
  float acc[15];  // gR(9) gT(3) ginvf li dl  -> 15
#pragma unroll
  for (int i = 0; i < 15; ++i) acc[i] = 0.f;
  float accS[AL_KMAX][14];  // dL/dS (12), dL/ds, dL/dt per incident edge slot
#pragma unroll
  for (int k = 0; k < AL_KMAX; ++k)
#pragma unroll
    for (int i = 0; i < 14; ++i) accS[k][i] = 0.f;

  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.HW; p += gridDim.x * blockDim.x) {
    const long long ip = (long long)n * a.HW + p;
    const float ld = a.logd[ip];
    const float d = __expf(ld);
    const float du = (float)(p % a.W) - cx, dv = (float)(p / a.W) - cy;
    const float xc = d * du * invf, yc = d * dv * invf, zc = d;
    const float Xw0 = R[0] * xc + R[1] * yc + R[2] * zc + R[3];
    const float Xw1 = R[4] * xc + R[5] * yc + R[6] * zc + R[7];
    const float Xw2 = R[8] * xc + R[9] * yc + R[10] * zc + R[11];
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, ginv = 0.f;
    const float inv = 1.0f / (d + 1e-6f);
#pragma unroll
    for (int k = 0; k < AL_KMAX; ++k) {
      if (k < ne) {
        const int e = a.edge_idx[e0 + k];
        const int g = e / a.group_size;
        const long long ep = (long long)e * a.HW + p;
        const float p0 = a.pred[3 * ep], p1 = a.pred[3 * ep + 1], p2 = a.pred[3 * ep + 2];
        const float* Sg = a.S + g * 12;
        const float Y0 = Sg[0] * p0 + Sg[1] * p1 + Sg[2] * p2 + Sg[3];
        const float Y1 = Sg[4] * p0 + Sg[5] * p1 + Sg[6] * p2 + Sg[7];
        const float Y2 = Sg[8] * p0 + Sg[9] * p1 + Sg[10] * p2 + Sg[11];
        const float r0 = Xw0 - Y0, r1 = Xw1 - Y1, r2 = Xw2 - Y2;
        const float nr = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
        const float w = fminf(a.weight[ep], 10.0f);
        acc[13] += w * nr;
        const float c = nr > 0.f ? w * invA / nr : 0.f;  // torch norm backward is 0 at the origin
        const float q0 = c * r0, q1 = c * r1, q2 = c * r2;
        g0 += q0; g1 += q1; g2 += q2;
        accS[k][0] -= q0 * p0; accS[k][1] -= q0 * p1; accS[k][2] -= q0 * p2; accS[k][3] -= q0;
        accS[k][4] -= q1 * p0; accS[k][5] -= q1 * p1; accS[k][6] -= q1 * p2; accS[k][7] -= q1;
        accS[k][8] -= q2 * p0; accS[k][9] -= q2 * p1; accS[k][10] -= q2 * p2; accS[k][11] -= q2;
        if (phaseB) {
          const float sg = a.st[g * 3], tg = a.st[g * 3 + 1], okg = a.st[g * 3 + 2];
          const float rho = a.invd[ep];
          const float m = (rho > 0.05f && okg != 0.f) ? 1.f : 0.f;
          const float res = inv - (sg * rho + tg);
          acc[14] += m * fabsf(res);
          const float sgn = (res > 0.f) ? 1.f : ((res < 0.f) ? -1.f : 0.f);
          const float gl = sgn * m * 2.0f * invA;
          ginv += gl;
          accS[k][12] -= gl * rho;
          accS[k][13] -= gl;
        }
      }
    }
    // gradient w.r.t. depth, then log-depth
    const float jx = R[0] * du * invf + R[1] * dv * invf + R[2];
    const float jy = R[4] * du * invf + R[5] * dv * invf + R[6];
    const float jz = R[8] * du * invf + R[9] * dv * invf + R[10];
    const float gd = g0 * jx + g1 * jy + g2 * jz - ginv * inv * inv;
    const float grad = gd * d;
    // torch.optim.Adam (no weight decay, no amsgrad): betas (0.9, 0.9), eps 1e-8
    const float m1 = 0.9f * a.adam_m[ip] + 0.1f * grad;
    const float v1 = 0.9f * a.adam_v[ip] + 0.1f * grad * grad;
    a.adam_m[ip] = m1;
    a.adam_v[ip] = v1;
    a.logd[ip] = ld - (lr / bc1) * m1 / (sqrtf(v1) / bc2s + 1e-8f);
    // pose / focal gradients
    acc[0] += g0 * xc; acc[1] += g0 * yc; acc[2] += g0 * zc;
    acc[3] += g1 * xc; acc[4] += g1 * yc; acc[5] += g1 * zc;
    acc[6] += g2 * xc; acc[7] += g2 * yc; acc[8] += g2 * zc;
    acc[9] += g0; acc[10] += g1; acc[11] += g2;
    acc[12] += d * (g0 * (R[0] * du + R[1] * dv) + g1 * (R[4] * du + R[5] * dv) + g2 * (R[8] * du + R[9] * dv));
  }
  // ---- reductions
  {
    float pose[12];
    pose[0] = acc[0]; pose[1] = acc[1]; pose[2] = acc[2]; pose[3] = acc[9];
    pose[4] = acc[3]; pose[5] = acc[4]; pose[6] = acc[5]; pose[7] = acc[10];
    pose[8] = acc[6]; pose[9] = acc[7]; pose[10] = acc[8]; pose[11] = acc[11];
    block_reduce_atomic<12>(pose, a.gpose + (long long)n * 12);
    float s3[3] = {acc[12], acc[13], acc[14]};
    block_reduce_atomic<3>(s3, a.gscal);
  }
#pragma unroll
  for (int k = 0; k < AL_KMAX; ++k) {
    if (k < ne) {  // uniform across the block
      const int g = a.edge_idx[e0 + k] / a.group_size;
      float s12[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) s12[i] = accS[k][i];
      block_reduce_atomic<12>(s12, a.gS + (long long)g * 12);
      float s2[2] = {accS[k][12], accS[k][13]};
      block_reduce_atomic<2>(s2, a.gst + (long long)g * 2);
    }
  }
}

// ---------------------------------------------------------------------------------------------- K11
// LAD scale/shift fit  min sum |s x + t - y|  with the reference's optimiser (Adam, default betas
// (0.9, 0.999), eps 1e-8; depth_eval.py:112-145), batched over G windows.  One kernel per iteration: every
// block reduces {sum sign(r) x, sum sign(r), sum |r|}; the last block of a window to finish applies Adam to
// (s, t), checks the |delta loss| < tol early exit and clears the accumulators.  state[g] = {s, t, m_s, v_s,
// m_t, v_t, prev_loss, step, done}.
__global__ void lad_step_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n_per_group,
                                float* __restrict__ state, double* __restrict__ acc, unsigned int* __restrict__ ticket,
                                float lr, float tol) {
  const int g = blockIdx.y;
  if (state[g * 9 + 8] != 0.f) return;  // converged (uniform per group: every block of the group skips)
  const float s = state[g * 9], t = state[g * 9 + 1];
  const float* xg = x + (long long)g * n_per_group;
  const float* yg = y + (long long)g * n_per_group;
  float a[3] = {0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_group;
       i += (long long)gridDim.x * blockDim.x) {
    const float xv = xg[i];
    const float r = s * xv + t - yg[i];
    const float sgn = (r > 0.f) ? 1.f : ((r < 0.f) ? -1.f : 0.f);
    a[0] += sgn * xv;
    a[1] += sgn;
    a[2] += fabsf(r);
  }
  block_reduce_atomic<3>(a, acc + g * 3);
  // the last block of this group to arrive applies the Adam update (no second launch, no host sync)
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(&ticket[g], 1u);
    if (tk == gridDim.x - 1) {
      __threadfence();
      volatile double* va = acc + g * 3;
      const float gs = (float)va[0], gt = (float)va[1], loss = (float)va[2];
      float* st = state + g * 9;
      const float step = st[7] + 1.f;
      const float bc1 = 1.f - powf(0.9f, step), bc2 = 1.f - powf(0.999f, step);
      st[2] = 0.9f * st[2] + 0.1f * gs;
      st[3] = 0.999f * st[3] + 0.001f * gs * gs;
      st[4] = 0.9f * st[4] + 0.1f * gt;
      st[5] = 0.999f * st[5] + 0.001f * gt * gt;
      st[0] -= (lr / bc1) * st[2] / (sqrtf(st[3]) / sqrtf(bc2) + 1e-8f);
      st[1] -= (lr / bc1) * st[4] / (sqrtf(st[5]) / sqrtf(bc2) + 1e-8f);
      if (st[7] > 0.f && fabsf(st[6] - loss) < tol) st[8] = 1.f;  // |delta loss| < tol: stop after this update
      st[6] = loss;
      st[7] = step;
      va[0] = 0.0; va[1] = 0.0; va[2] = 0.0;
      ticket[g] = 0u;
      __threadfence();
    }
  }
}

// Whole LAD fit in ONE cooperative launch: every block keeps its slice of (x, y) in shared memory (when it fits)
// and runs up to `iters` Adam iterations with the current (s, t) in registers.  Per iteration a block reduces its
// {sum sign(r) x, sum sign(r), sum |r|} in a fixed tree, stores the three sums as its row of `part` and takes a ticket;
// the LAST block of the window folds the rows in block order (no atomics: the fit is bit-reproducible), applies Adam
// and PUBLISHES the new (s, t, done) as two 8-byte words that each carry the iteration number -- the other blocks
// spin on those words only (one L2 round trip from the update to the next iteration, no separate state reads).
// Same arithmetic and early exit as lad_step_kernel.  Workspace per window (doubles, geo4d_lad_fit_workspace_doubles):
// [0] {s, iteration} [1] {t, iteration | done << 31} [2] ticket [3 ...] part[blocks][3] floats.
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(512, 1)
lad_fit_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n_per_group,
               float* __restrict__ state, double* __restrict__ ws, int ws_stride, float lr, float tol, int iters,
               int cache) {
  extern __shared__ float lsm[];   // [2][slice] x | y when cache != 0
  __shared__ float s_red[16][3];
  __shared__ float s_st[3];        // s, t, done for the next iteration
  __shared__ int s_last;
  const int g = blockIdx.y;
  const int nb = gridDim.x;
  const long long per = (n_per_group + nb - 1) / nb;
  const long long i0 = (long long)blockIdx.x * per;
  const long long i1 = i0 + per < n_per_group ? i0 + per : n_per_group;
  const int cnt = (int)(i1 > i0 ? i1 - i0 : 0);
  const float* xg = x + (long long)g * n_per_group + i0;
  const float* yg = y + (long long)g * n_per_group + i0;
  unsigned long long* pub = reinterpret_cast<unsigned long long*>(ws + (size_t)g * ws_stride);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(pub + 2);
  float* part = reinterpret_cast<float*>(pub + 3);
  volatile float* vst = state + g * 9;
  if (cache) {
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) { lsm[i] = xg[i]; lsm[per + i] = yg[i]; }
  }
  float sc = vst[0], tc = vst[1];
  if (vst[8] != 0.f) return;   // already converged (uniform over the blocks of this window)
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    float a[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float xv = cache ? lsm[i] : xg[i];
      const float yv = cache ? lsm[per + i] : yg[i];
      const float r = sc * xv + tc - yv;
      const float sgn = (r > 0.f) ? 1.f : ((r < 0.f) ? -1.f : 0.f);
      a[0] += sgn * xv;
      a[1] += sgn;
      a[2] += fabsf(r);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = warp_sum(a[k]);
      if (lane == 0) s_red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float t3[3] = {0.f, 0.f, 0.f};
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t3[0] += s_red[w][0]; t3[1] += s_red[w][1]; t3[2] += s_red[w][2]; }
      __stcg(part + blockIdx.x * 3 + 0, t3[0]); __stcg(part + blockIdx.x * 3 + 1, t3[1]); __stcg(part + blockIdx.x * 3 + 2, t3[2]);
      __threadfence();
      s_last = (atomicAdd(ticket, 1u) == (unsigned)nb - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (warp == 0) {
        // fold the rows in a fixed order: lane l takes blocks l, l+32, ... then a shuffle tree
        double d[3] = {0.0, 0.0, 0.0};
        for (int b = lane; b < nb; b += 32) {
          d[0] += (double)__ldcg(part + b * 3 + 0); d[1] += (double)__ldcg(part + b * 3 + 1); d[2] += (double)__ldcg(part + b * 3 + 2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          d[0] += __shfl_xor_sync(0xffffffffu, d[0], o); d[1] += __shfl_xor_sync(0xffffffffu, d[1], o);
          d[2] += __shfl_xor_sync(0xffffffffu, d[2], o);
        }
        if (lane == 0) {
          const float gs = (float)d[0], gt = (float)d[1], loss = (float)d[2];
          const float step = vst[7] + 1.f;
          const float bc1 = 1.f - powf(0.9f, step), bc2 = 1.f - powf(0.999f, step);
          const float ms = 0.9f * vst[2] + 0.1f * gs, vs = 0.999f * vst[3] + 0.001f * gs * gs;
          const float mt = 0.9f * vst[4] + 0.1f * gt, vt = 0.999f * vst[5] + 0.001f * gt * gt;
          const float sn = vst[0] - (lr / bc1) * ms / (sqrtf(vs) / sqrtf(bc2) + 1e-8f);
          const float tn = vst[1] - (lr / bc1) * mt / (sqrtf(vt) / sqrtf(bc2) + 1e-8f);
          const float done = (vst[7] > 0.f && fabsf(vst[6] - loss) < tol) ? 1.f : 0.f;   // |delta loss| < tol: stop after this update
          vst[2] = ms; vst[3] = vs; vst[4] = mt; vst[5] = vt;
          vst[0] = sn; vst[1] = tn; vst[6] = loss; vst[7] = step; vst[8] = done;
          *ticket = 0u;
          __threadfence();
          const unsigned int gen = (unsigned)(it + 1);
          st_volatile_u64(pub + 0, ((unsigned long long)gen << 32) | __float_as_uint(sn));
          st_volatile_u64(pub + 1, ((unsigned long long)(gen | (done != 0.f ? 0x80000000u : 0u)) << 32) | __float_as_uint(tn));
          s_st[0] = sn; s_st[1] = tn; s_st[2] = done;
        }
      }
    } else if (threadIdx.x == 0) {
      const unsigned int gen = (unsigned)(it + 1);
      unsigned long long w0, w1;
      long long t0 = clock64();
      while ((unsigned)((w0 = ld_volatile_u64(pub + 0)) >> 32) != gen ||
             ((unsigned)((w1 = ld_volatile_u64(pub + 1)) >> 32) & 0x7fffffffu) != gen) {
        if (clock64() - t0 > G4_MBAR_TIMEOUT_CYCLES) {
          printf("g4: lad_fit grid barrier timeout (block %d,%d it %d)\n", (int)blockIdx.x, (int)blockIdx.y, it);
          __trap();
        }
      }
      s_st[0] = __uint_as_float((unsigned)w0);
      s_st[1] = __uint_as_float((unsigned)w1);
      s_st[2] = ((unsigned)(w1 >> 32) & 0x80000000u) ? 1.f : 0.f;
    }
    __syncthreads();
    sc = s_st[0]; tc = s_st[1];
    const bool stop = s_st[2] != 0.f;
    __syncthreads();   // s_st / s_red / s_last are rewritten next iteration
    if (stop) break;
  }
}

// delta < 1.25 accuracy of s*x + t against y under mask (depth_eval.py:296-317): out[g] = {count_ok, count}
__global__ void delta125_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                const float* __restrict__ w, long long n_per_group, const float* __restrict__ st,
                                int st_stride, double* __restrict__ out) {
  const int g = blockIdx.y;
  const float s = st[g * st_stride], t = st[g * st_stride + 1];
  float a[2] = {0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_group;
       i += (long long)gridDim.x * blockDim.x) {
    const long long j = (long long)g * n_per_group + i;
    const float xv = x[j], gt = y[j];
    if (w[j] > 0.5f && xv > 0.05f && gt > 0.f) {
      const float pr = fmaxf(s * xv + t, 1e-5f);
      const float ratio = fmaxf(pr / gt, gt / pr);
      a[0] += (ratio < 1.25f) ? 1.f : 0.f;
      a[1] += 1.f;
    }
  }
  block_reduce_atomic<2>(a, out + g * 2);
}


// ---------------------------------------------------------------------------------------------- PnP moments
// Reductions of the SQPnP normal equations (Terzakis & Lourakis 2020, the solver behind
// cv2.solvePnPRansac(flags=SOLVEPNP_SQPNP) as called by fast_pnp, init_im_poses.py:824-865).  With pixel
// offsets (du, dv) = (u - cx, v - cy), r2 = du^2 + dv^2 and world point m, every entry of sum Q, sum Q A and
// sum A^T Q A is (1/f)^k times one of
//   out[0..4)   = { n, sum du, sum dv, sum r2 }
//   out[4..16)  = { sum m, sum du m, sum dv m, sum r2 m }                       (3 each)
//   out[16..40) = { sum m m^T, sum du m m^T, sum dv m m^T, sum r2 m m^T }        (6 unique each: xx xy xz yy yz zz)
//   out[40]     = number of points used
// so ONE pass serves every tentative focal.  With `gate` (per (frame, candidate) world-to-camera [R|t] and f)
// only points with conf > 0.5, positive depth and reprojection error < thr pixels contribute: this is the
// consensus set of the RANSAC stage, re-fitted by the same solver (what OpenCV does after RANSAC).
__global__ void __launch_bounds__(256)
pnp_moments_kernel(const float* __restrict__ pts, const float* __restrict__ conf, int HW, int W, float cx, float cy,
                   const float* __restrict__ gate /*[F][C][13] or null*/, int C, float thr2,
                   double* __restrict__ out /*[F][C][41]*/) {
  const int f = blockIdx.y, c = blockIdx.z;
  float acc[41];
#pragma unroll
  for (int i = 0; i < 41; ++i) acc[i] = 0.f;
  float G[13];
  if (gate) {
#pragma unroll
    for (int i = 0; i < 13; ++i) G[i] = gate[((long long)f * C + c) * 13 + i];
  }
  const float* P = pts + (long long)f * HW * 3;
  const float* Cf = conf + (long long)f * HW;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    if (!(Cf[p] > 0.5f)) continue;
    const float X = P[3 * p], Y = P[3 * p + 1], Z = P[3 * p + 2];
    const float u = (float)(p % W), v = (float)(p / W);
    if (gate) {
      const float xc = G[0] * X + G[1] * Y + G[2] * Z + G[3];
      const float yc = G[4] * X + G[5] * Y + G[6] * Z + G[7];
      const float zc = G[8] * X + G[9] * Y + G[10] * Z + G[11];
      if (!(zc > 0.f)) continue;
      const float eu = u - (G[12] * xc / zc + cx), ev = v - (G[12] * yc / zc + cy);
      if (!(eu * eu + ev * ev < thr2)) continue;
    }
    const float du = u - cx, dv = v - cy, r2 = du * du + dv * dv;
    const float mm[6] = {X * X, X * Y, X * Z, Y * Y, Y * Z, Z * Z};
    acc[0] += 1.f; acc[1] += du; acc[2] += dv; acc[3] += r2;
    acc[4] += X; acc[5] += Y; acc[6] += Z;
    acc[7] += du * X; acc[8] += du * Y; acc[9] += du * Z;
    acc[10] += dv * X; acc[11] += dv * Y; acc[12] += dv * Z;
    acc[13] += r2 * X; acc[14] += r2 * Y; acc[15] += r2 * Z;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      acc[16 + k] += mm[k];
      acc[22 + k] += du * mm[k];
      acc[28 + k] += dv * mm[k];
      acc[34 + k] += r2 * mm[k];
    }
    acc[40] += 1.f;
  }
  block_reduce_atomic<41>(acc, out + ((long long)f * C + c) * 41);
}

// Shift / focal fit of point_map_to_depth (utils/geometry.py:162-270): for a trial z-shift s per window the
// reference minimises sum |f p - uv|^2 with p = xy / (z + s) and the optimal f = S1 / S2, S1 = sum p.uv,
// S2 = sum |p|^2.  out[g] = { S1, S2, dS1/ds, dS2/ds, sum |uv|^2, n } over pixels with conf > 0.5.
__global__ void shift_focal_sums_kernel(const float* __restrict__ pts, const float* __restrict__ conf, int HW, int W,
                                        int H, const float* __restrict__ shift, float zoff,
                                        double* __restrict__ out) {
  const int g = blockIdx.y;
  const float s = shift[g];
  const float ar = (float)W / (float)H;
  const float sx = ar / sqrtf(1.f + ar * ar), sy = 1.f / sqrtf(1.f + ar * ar);
  float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* P = pts + (long long)g * HW * 3;
  const float* Cf = conf + (long long)g * HW;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    if (!(Cf[p] > 0.5f)) continue;
    // image_plane_uv: linspace(-span*(n-1)/n, span*(n-1)/n, n)
    const int iu = p % W, iv = p / W;
    const float uu = W > 1 ? sx * (float)(W - 1) / (float)W * (2.f * (float)iu / (float)(W - 1) - 1.f) : 0.f;
    const float vv = H > 1 ? sy * (float)(H - 1) / (float)H * (2.f * (float)iv / (float)(H - 1) - 1.f) : 0.f;
    const float X = P[3 * p], Y = P[3 * p + 1], Z = P[3 * p + 2] + zoff;
    const float iz = 1.0f / (Z + s);
    const float px = X * iz, py = Y * iz;
    const float dot = px * uu + py * vv, n2 = px * px + py * py;
    a[0] += dot;
    a[1] += n2;
    a[2] -= dot * iz;        // d(p.uv)/ds
    a[3] -= 2.f * n2 * iz;   // d|p|^2/ds
    a[4] += uu * uu + vv * vv;
    a[5] += 1.f;
  }
  block_reduce_atomic<6>(a, out + g * 6);
}

// ---------------------------------------------------------------------------------------------- small parameters
}  // namespace g4
#include "align_small.cuh"
namespace g4 {

__global__ void __launch_bounds__(256)
align_small_kernel(const SmallArgs a) {
  extern __shared__ float sh[];
  align_small_body(a, a.it ? *a.it : 0, sh);
}

int device_sm_count();

static inline int blocks_for(long long n, int cap) {
  long long b = (n + 255) / 256;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace g4

using namespace g4;
#define G4_STREAM cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_)

extern "C" int geo4d_postprocess_window(const float* maps, int64_t thw, float* pts, float* inv_conf, float* invdepth,
                                        unsigned char* valid, float sky_value, float sky_eps, float far_value,
                                        float alpha, float beta, int has_conf, g4_stream_t stream_) {
  G4_STREAM;
  if (!maps || !pts || !inv_conf || !invdepth) { set_last_error("postprocess_window: null"); return G4_ERR_BAD_ARG; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  postprocess_window_kernel<<<blocks_for(thw, sms * 8), 256, 0, stream>>>(maps, thw, pts, inv_conf, invdepth, valid,
                                                                         sky_value, sky_eps, far_value, alpha, beta,
                                                                         has_conf);
  return check_launch("postprocess_window");
}

extern "C" int geo4d_raymap_moments(const float* raydir, const float* raymoment, int T, int H, int W, double* out,
                                    g4_stream_t stream_) {
  G4_STREAM;
  if (!raydir || !raymoment || !out || T < 1 || T > 65535) { set_last_error("raymap_moments: bad args"); return G4_ERR_BAD_ARG; }
  const int S = H < W ? H : W;
  const int x0 = (W - S) / 2, y0 = (H - S) / 2;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * 18 * T, stream);
  if (e != cudaSuccess) { set_last_error("raymap_moments: memset: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  dim3 grid(blocks_for((long long)S * S, 32), T);
  raymap_moments_kernel<<<grid, 256, 0, stream>>>(raydir, raymoment, T, H, W, x0, y0, S, out);
  return check_launch("raymap_moments");
}

extern "C" int geo4d_transform_points(const float* x, int n_sets, int64_t pts_per_set, const float* mats, float* out,
                                      int mode, g4_stream_t stream_) {
  G4_STREAM;
  if (!x || !mats || !out || n_sets < 1 || n_sets > 65535 || pts_per_set < 1 || (mode != 0 && mode != 1)) {
    set_last_error("transform_points: bad args"); return G4_ERR_BAD_ARG;
  }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  int bx = (4 * sms + n_sets - 1) / n_sets;
  const long long maxbx = (pts_per_set + 255) / 256;
  if (bx > maxbx) bx = (int)maxbx;
  if (bx < 1) bx = 1;
  transform_points_kernel<<<dim3(bx, n_sets), 256, 0, stream>>>(x, pts_per_set, mats, out, mode);
  return check_launch("transform_points");
}

extern "C" int geo4d_umeyama_moments(const float* x, const float* y, const float* w1, const float* w2, int64_t n,
                                     int pass, const double* means, double* out, g4_stream_t stream_) {
  G4_STREAM;
  if (!x || !y || !w1 || !out || (pass == 1 && !means) || pass < 0 || pass > 1) {
    set_last_error("umeyama_moments: bad args"); return G4_ERR_BAD_ARG;
  }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * 10, stream);
  if (e != cudaSuccess) { set_last_error("umeyama_moments: memset: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  umeyama_moments_kernel<<<blocks_for(n, sms * 4), 256, 0, stream>>>(x, y, w1, w2, n, pass, means, out);
  return check_launch("umeyama_moments");
}

extern "C" int geo4d_align_iter(float* logd, float* adam_m, float* adam_v, const float* pred, const float* weight,
                                const float* invd, const int* edge_ptr, const int* edge_idx, const float* poses,
                                const float* S, const float* scal, const float* invf, const int* it, const float* st,
                                double* gpose, double* gS, double* gscal, double* gst, int N, int G, int HW, int W,
                                int group_size, int max_edges_per_image, g4_stream_t stream_) {
  G4_STREAM;
  if (!logd || !adam_m || !adam_v || !pred || !weight || !edge_ptr || !edge_idx || !poses || !S || !scal || !invf || !st ||
      !gpose || !gS || !gscal || !gst) { set_last_error("align_iter: null pointer"); return G4_ERR_BAD_ARG; }
  if (max_edges_per_image > AL_KMAX) {
    set_last_error("align_iter: an image is observed by %d windows; at most %d supported", max_edges_per_image, AL_KMAX);
    return G4_ERR_UNSUPPORTED;
  }
  if (N < 1 || N > 65535) { set_last_error("align_iter: N=%d", N); return G4_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(gpose, 0, sizeof(double) * 12 * N, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(gS, 0, sizeof(double) * 12 * G, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(gscal, 0, sizeof(double) * 3, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(gst, 0, sizeof(double) * 2 * G, stream);
  if (e != cudaSuccess) { set_last_error("align_iter: memset: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  AlignIterArgs a;
  a.logd = logd; a.adam_m = adam_m; a.adam_v = adam_v; a.pred = pred; a.weight = weight; a.invd = invd;
  a.edge_ptr = edge_ptr; a.edge_idx = edge_idx; a.poses = poses; a.S = S; a.scal = scal; a.invf = invf; a.it = it; a.st = st;
  a.gpose = gpose; a.gS = gS; a.gscal = gscal; a.gst = gst;
  a.N = N; a.HW = HW; a.W = W; a.group_size = group_size;
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  int bx = (4 * sms + N - 1) / N;
  const int maxbx = (HW + 255) / 256;
  if (bx > maxbx) bx = maxbx;
  if (bx < 1) bx = 1;
  dim3 grid(bx, N);
  align_iter_kernel<<<grid, 256, 0, stream>>>(a);
  return check_launch("align_iter");
}

extern "C" int geo4d_lad_step(const float* x, const float* y, int64_t n_per_group, int G, float* state, double* acc,
                              float lr, float tol, g4_stream_t stream_) {
  G4_STREAM;
  if (!x || !y || !state || !acc || G < 1 || G > 65535) { set_last_error("lad_step: bad args"); return G4_ERR_BAD_ARG; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  int bx = (2 * sms + G - 1) / G;
  if (bx < 1) bx = 1;
  dim3 grid(bx, G);
  // acc holds 3 doubles per window followed by one 32-bit arrival ticket per window (all zero initially)
  unsigned int* ticket = reinterpret_cast<unsigned int*>(acc + 3 * (size_t)G);
  lad_step_kernel<<<grid, 256, 0, stream>>>(x, y, n_per_group, state, acc, ticket, lr, tol);
  return check_launch("lad_step");
}

extern "C" size_t geo4d_lad_fit_workspace_doubles(int G) {
  const int sms = device_sm_count();
  const int nb = sms > 0 ? sms : 256;
  return (size_t)(G < 1 ? 1 : G) * (size_t)(3 + (3 * nb + 1) / 2);
}

extern "C" int geo4d_lad_fit(const float* x, const float* y, int64_t n_per_group, int G, float* state, double* acc,
                             float lr, float tol, int iters, g4_stream_t stream_) {
  G4_STREAM;
  if (!x || !y || !state || !acc || G < 1 || G > 65535 || iters < 0) { set_last_error("lad_fit: bad args"); return G4_ERR_BAD_ARG; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  if (G > sms) { set_last_error("lad_fit: %d windows > %d SMs (use geo4d_lad_step)", G, sms); return G4_ERR_UNSUPPORTED; }
  int bx = sms / G;   // one 512-thread block per SM: all blocks are co-resident (cooperative launch checks it)
  if ((long long)bx > (n_per_group + 511) / 512) bx = (int)((n_per_group + 511) / 512);
  if (bx < 1) bx = 1;
  const long long per = (n_per_group + bx - 1) / bx;
  size_t smem = (size_t)per * 2 * sizeof(float);
  int cache = 1;
  if (smem > 200 * 1024) { smem = 0; cache = 0; }
  static unsigned long long attr_mask = 0;
  if (once_per_device(&attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(lad_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { unlatch_device(&attr_mask); set_last_error("lad_fit: smem attr: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  }
  // workspace (geo4d_lad_fit_workspace_doubles(G) doubles): per window {s | it}, {t | it | done}, ticket, part[blocks][3]
  const int ws_stride = 3 + (3 * sms + 1) / 2;
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * (size_t)G * ws_stride, stream);
  if (e != cudaSuccess) { set_last_error("lad_fit: memset: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(bx, G); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, lad_fit_kernel, x, y, (long long)n_per_group, state, acc, ws_stride, lr, tol, iters, cache);
  if (e != cudaSuccess) { set_last_error("lad_fit: launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return G4_ERR_CUDA; }
  return check_launch("lad_fit");
}

extern "C" int geo4d_delta125(const float* x, const float* y, const float* w, int64_t n_per_group, int G,
                              const float* st, int st_stride, double* out, g4_stream_t stream_) {
  G4_STREAM;
  if (!x || !y || !w || !st || !out || G < 1 || G > 65535) { set_last_error("delta125: bad args"); return G4_ERR_BAD_ARG; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * 2 * G, stream);
  if (e != cudaSuccess) { set_last_error("delta125: memset: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  int bx = (2 * sms + G - 1) / G;
  if (bx < 1) bx = 1;
  dim3 grid(bx, G);
  delta125_kernel<<<grid, 256, 0, stream>>>(x, y, w, n_per_group, st, st_stride, out);
  return check_launch("delta125");
}

extern "C" int geo4d_pnp_moments(const float* pts, const float* conf, int F, int HW, int W, float cx, float cy,
                                 const float* gate, int C, float thr_px, double* out, g4_stream_t stream_) {
  G4_STREAM;
  if (!pts || !conf || !out || F < 1 || F > 65535 || C < 1 || C > 65535) { set_last_error("pnp_moments: bad args"); return G4_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * 41 * F * C, stream);
  if (e != cudaSuccess) { set_last_error("pnp_moments: memset: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  int bx = (2 * sms + F * C - 1) / (F * C);
  const int maxbx = (HW + 255) / 256;
  if (bx > maxbx) bx = maxbx;
  if (bx < 1) bx = 1;
  dim3 grid(bx, F, C);
  pnp_moments_kernel<<<grid, 256, 0, stream>>>(pts, conf, HW, W, cx, cy, gate, C, thr_px * thr_px, out);
  return check_launch("pnp_moments");
}

extern "C" int geo4d_shift_focal_sums(const float* pts, const float* conf, int G, int HW, int W, int H,
                                      const float* shift, float zoff, double* out, g4_stream_t stream_) {
  G4_STREAM;
  if (!pts || !conf || !shift || !out || G < 1 || G > 65535) { set_last_error("shift_focal_sums: bad args"); return G4_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * 6 * G, stream);
  if (e != cudaSuccess) { set_last_error("shift_focal_sums: memset: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  int bx = (2 * sms + G - 1) / G;
  const int maxbx = (HW + 255) / 256;
  if (bx > maxbx) bx = maxbx;
  if (bx < 1) bx = 1;
  dim3 grid(bx, G);
  shift_focal_sums_kernel<<<grid, 256, 0, stream>>>(pts, conf, HW, W, H, shift, zoff, out);
  return check_launch("shift_focal_sums");
}

extern "C" size_t geo4d_align_small_adam_floats(int N, int G) { return 2 * ((size_t)N * 7 + 1 + (size_t)G * 8 + G + G + (size_t)G * 8); }

extern "C" int geo4d_align_small_step(float* im_poses, float* im_focal, float* pw_poses, float* s_depth, float* t_depth,
                                      float* ta_poses, float* adam, const double* gpose, const double* gS,
                                      const double* gscal, const double* gst, const float* traj, const int* e_img,
                                      const int* edge_ptr, const int* edge_idx, const float* valid_traj,
                                      const float* scal, const int* it, float* poses_out, float* S_out,
                                      float* invf_out, float* st_out, int N, int G, int group_size, int start_b,
                                      float temporal_smoothing_weight, float translation_weight, float base_scale,
                                      float focal_break, g4_stream_t stream_) {
  G4_STREAM;
  if (!im_poses || !im_focal || !pw_poses || !s_depth || !t_depth || !ta_poses || !adam || !gpose || !gS || !gscal ||
      !gst || !traj || !e_img || !edge_ptr || !edge_idx || !valid_traj || !scal || !poses_out || !S_out || !invf_out ||
      !st_out) { set_last_error("align_small_step: null pointer"); return G4_ERR_BAD_ARG; }
  SmallArgs a;
  a.im_poses = im_poses; a.im_focal = im_focal; a.pw_poses = pw_poses; a.s_depth = s_depth; a.t_depth = t_depth;
  a.ta_poses = ta_poses; a.adam = adam; a.gpose = gpose; a.gS = gS; a.gscal = gscal; a.gst = gst; a.traj = traj;
  a.e_img = e_img; a.edge_ptr = edge_ptr; a.edge_idx = edge_idx; a.valid_traj = valid_traj; a.scal = scal; a.it = it;
  a.poses_out = poses_out; a.S_out = S_out; a.invf_out = invf_out; a.st_out = st_out;
  a.N = N; a.G = G; a.gs = group_size; a.start_b = start_b;
  a.tsw = temporal_smoothing_weight; a.tw = translation_weight; a.log_base_scale = logf(base_scale);
  a.focal_break = focal_break;
  const size_t smem = sizeof(float) * align_small_smem_floats(N, G);
  if (smem > 200 * 1024) { set_last_error("align_small_step: too many images/windows for one CTA (%d, %d)", N, G); return G4_ERR_UNSUPPORTED; }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(align_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error("align_small_step: smem attr: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  }
  align_small_kernel<<<1, 256, smem, stream>>>(a);
  return check_launch("align_small_step");
}
