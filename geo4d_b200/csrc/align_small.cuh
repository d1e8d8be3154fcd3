// O(N + G) part of one alignment step (shared by the stand-alone geo4d_align_small_step kernel and by the
// persistent alignment loop in align_loop.cu).
#pragma once
#include "common.cuh"

namespace g4 {

// ---------------------------------------------------------------------------------------------- small parameters
// Everything of one optimisation step that is O(N + G): chain rule from the matrix-form gradients reduced by
// align_iter_kernel to the reference's parametrisations (unit quaternion xyzw + signed-log translation,
// base_opt_group.py:260-288; window scale exp(p7) with the mean-normalisation of :303-314; focal exp(p/20),
// optimizer_group.py:193-198), the two pose-graph terms (temporal smoothing and trajectory prior,
// optimizer_group.py:492-519 with relative_pose_loss :529-542), torch.optim.Adam (betas 0.9/0.9, eps 1e-8) for
// every small parameter, and the refreshed pose / sim(3) / focal matrices the next dense iteration reads.
// One CTA; parameters live in registers / shared memory.
struct SmallArgs {
  float* im_poses;   // [N][7]
  float* im_focal;   // [1]
  float* pw_poses;   // [G][8]
  float* s_depth;    // [G]
  float* t_depth;    // [G]
  float* ta_poses;   // [G][8]
  float* adam;       // m then v for each tensor above, in the same order
  const double* gpose; const double* gS; const double* gscal; const double* gst;
  const float* traj;        // [G*gs][16] c2w rows
  const int* e_img;         // [G*gs] image of each (window, frame)
  const int* edge_ptr; const int* edge_idx;
  const float* valid_traj;  // [G]
  const float* scal; const int* it;
  float* poses_out; float* S_out; float* invf_out; float* st_out;
  int N, G, gs, start_b;
  float tsw, tw, log_base_scale, focal_break;
};

__device__ __forceinline__ void quat_to_R(const float* q4, float* R, float* qn, float* inv_norm) {
  const float nrm = sqrtf(q4[0] * q4[0] + q4[1] * q4[1] + q4[2] * q4[2] + q4[3] * q4[3]);
  const float in = 1.0f / nrm;
  const float x = q4[0] * in, y = q4[1] * in, z = q4[2] * in, w = q4[3] * in;
  qn[0] = x; qn[1] = y; qn[2] = z; qn[3] = w; *inv_norm = in;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ float sexpm1(float v) { return v >= 0.f ? expm1f(v) : -expm1f(-v); }

// gradient of the loss w.r.t. the raw quaternion given dL/dR (row-major 3x3)
__device__ __forceinline__ void quat_backward(const float* qn, float inv_norm, const float* G, float* gq) {
  const float x = qn[0], y = qn[1], z = qn[2], w = qn[3];
  float g[4];
  g[0] = 2 * y * (G[1] + G[3]) + 2 * z * (G[2] + G[6]) - 4 * x * (G[4] + G[8]) + 2 * w * (G[7] - G[5]);
  g[1] = 2 * x * (G[1] + G[3]) - 4 * y * (G[0] + G[8]) + 2 * z * (G[5] + G[7]) + 2 * w * (G[2] - G[6]);
  g[2] = 2 * x * (G[2] + G[6]) + 2 * y * (G[5] + G[7]) - 4 * z * (G[0] + G[4]) + 2 * w * (G[3] - G[1]);
  g[3] = 2 * z * (G[3] - G[1]) + 2 * y * (G[2] - G[6]) + 2 * x * (G[7] - G[5]);
  const float dot = x * g[0] + y * g[1] + z * g[2] + w * g[3];
  gq[0] = (g[0] - x * dot) * inv_norm; gq[1] = (g[1] - y * dot) * inv_norm;
  gq[2] = (g[2] - z * dot) * inv_norm; gq[3] = (g[3] - w * dot) * inv_norm;
}

// l(A, B) = ||A_R^T B_R - I||_F + tw ||A_R^T (B_T - A_T)||; accumulate wgt * dl into (gRA, gtA) and/or (gRB, gtB)
__device__ __forceinline__ void relpose_grad(const float* RA, const float* tA, const float* RB, const float* tB,
                                             float tw, float wgt, float* gRA, float* gtA, float* gRB, float* gtB) {
  float M[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      M[i * 3 + j] = RA[0 + i] * RB[0 + j] + RA[3 + i] * RB[3 + j] + RA[6 + i] * RB[6 + j] - (i == j ? 1.f : 0.f);
  float n2 = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) n2 += M[i] * M[i];
  const float inr = n2 > 0.f ? wgt / sqrtf(n2) : 0.f;
  const float dt[3] = {tB[0] - tA[0], tB[1] - tA[1], tB[2] - tA[2]};
  float d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = RA[0 + i] * dt[0] + RA[3 + i] * dt[1] + RA[6 + i] * dt[2];
  const float m2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const float im = m2 > 0.f ? wgt * tw / sqrtf(m2) : 0.f;
  const float gd[3] = {d[0] * im, d[1] * im, d[2] * im};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float rg = RA[k * 3 + 0] * gd[0] + RA[k * 3 + 1] * gd[1] + RA[k * 3 + 2] * gd[2];
    if (gtB) gtB[k] += rg;
    if (gtA) gtA[k] -= rg;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (gRA) gRA[k * 3 + i] += inr * (RB[k * 3 + 0] * M[i * 3 + 0] + RB[k * 3 + 1] * M[i * 3 + 1] + RB[k * 3 + 2] * M[i * 3 + 2]) + dt[k] * gd[i];
      if (gRB) gRB[k * 3 + i] += inr * (RA[k * 3 + 0] * M[0 * 3 + i] + RA[k * 3 + 1] * M[1 * 3 + i] + RA[k * 3 + 2] * M[2 * 3 + i]);
    }
  }
}

__device__ __forceinline__ float adam_step(float p, float g, float* m, float* v, float lr, float bc1, float bc2s) {
  const float m1 = 0.9f * (*m) + 0.1f * g;
  const float v1 = 0.9f * (*v) + 0.1f * g * g;
  *m = m1; *v = v1;
  return p - (lr / bc1) * m1 / (sqrtf(v1) / bc2s + 1e-8f);
}

// The body runs on ONE CTA of any size (threadIdx.x / blockDim.x strides, __syncthreads inside); `sh` needs
// align_small_smem_floats(N, G) floats.  `it` = iteration whose `scal` row / bias corrections apply.
__host__ __device__ inline size_t align_small_smem_floats(int N, int G) { return (size_t)N * 17 + (size_t)G * 37; }

__device__ __forceinline__ void align_small_body(const SmallArgs& a, const int it, float* sh) {
  const int N = a.N, G = a.G, gs = a.gs;
  float* sP = sh;                 // [N][12] current image poses (R rows | T as [R00 R01 R02 T0 ...])
  float* sQ = sP + N * 12;        // [N][5] normalised quaternion + 1/|q|
  float* sW = sQ + N * 5;         // [G][12] window [R | T'] (unscaled)
  float* sWq = sW + G * 12;       // [G][6]  qn(4), 1/|q|, s_g
  float* sA = sWq + G * 6;        // [G][12] trajectory-alignment [Ra | Ta]
  float* sAq = sA + G * 12;       // [G][6]  qn(4), 1/|q|, sigma
  float* sRed = sAq + G * 6;      // [G] dL/ds_g * s_g, then scratch
  const float* sc = a.scal + (long long)it * 8;
  const float lr = sc[3];
  const bool phaseB = sc[7] != 0.f;
  const float stepA = (float)(it + 1), stepB = (float)(it - a.start_b + 1);
  const float bc1A = 1.f - powf(0.9f, stepA), bc2A = sqrtf(1.f - powf(0.9f, stepA));
  const float bc1B = phaseB ? 1.f - powf(0.9f, stepB) : 1.f, bc2B = phaseB ? sqrtf(1.f - powf(0.9f, stepB)) : 1.f;
  // Adam state offsets
  float* mP = a.adam;              float* vP = mP + N * 7;
  float* mF = vP + N * 7;          float* vF = mF + 1;
  float* mW = vF + 1;              float* vW = mW + G * 8;
  float* mS = vW + G * 8;          float* vS = mS + G;
  float* mT = vS + G;              float* vT = mT + G;
  float* mA = vT + G;              float* vA = mA + G * 8;

  // ---- phase 0: matrices from the current parameters
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float R[9], qn[4], in;
    quat_to_R(a.im_poses + n * 7, R, qn, &in);
    float* P = sP + n * 12;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      P[i * 4 + 0] = R[i * 3 + 0]; P[i * 4 + 1] = R[i * 3 + 1]; P[i * 4 + 2] = R[i * 3 + 2];
      P[i * 4 + 3] = sexpm1(a.im_poses[n * 7 + 4 + i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) sQ[n * 5 + i] = qn[i];
    sQ[n * 5 + 4] = in;
  }
  float mean7 = 0.f;
  for (int g = 0; g < G; ++g) mean7 += a.pw_poses[g * 8 + 7];
  mean7 /= (float)G;
  const float normF = expf(a.log_base_scale - mean7);
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float R[9], qn[4], in;
    quat_to_R(a.pw_poses + g * 8, R, qn, &in);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      sW[g * 12 + i * 4 + 0] = R[i * 3 + 0]; sW[g * 12 + i * 4 + 1] = R[i * 3 + 1]; sW[g * 12 + i * 4 + 2] = R[i * 3 + 2];
      sW[g * 12 + i * 4 + 3] = sexpm1(a.pw_poses[g * 8 + 4 + i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) sWq[g * 6 + i] = qn[i];
    sWq[g * 6 + 4] = in;
    sWq[g * 6 + 5] = expf(a.pw_poses[g * 8 + 7]) * normF;
    quat_to_R(a.ta_poses + g * 8, R, qn, &in);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      sA[g * 12 + i * 4 + 0] = R[i * 3 + 0]; sA[g * 12 + i * 4 + 1] = R[i * 3 + 1]; sA[g * 12 + i * 4 + 2] = R[i * 3 + 2];
      sA[g * 12 + i * 4 + 3] = sexpm1(a.ta_poses[g * 8 + 4 + i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) sAq[g * 6 + i] = qn[i];
    sAq[g * 6 + 4] = in;
    sAq[g * 6 + 5] = expf(a.ta_poses[g * 8 + 7]);
  }
  __syncthreads();

  auto load_RT = [](const float* P, float* R, float* t) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { R[i * 3] = P[i * 4]; R[i * 3 + 1] = P[i * 4 + 1]; R[i * 3 + 2] = P[i * 4 + 2]; t[i] = P[i * 4 + 3]; }
  };
  auto traj_pose = [&](int e, int g, float* YR, float* Yt) {  // Y = [Ra | Ta] [R_tr | sigma T_tr]
    const float* Tr = a.traj + (long long)e * 16;
    float Ra[9], Ta[3];
    load_RT(sA + g * 12, Ra, Ta);
    const float sg = sAq[g * 6 + 5];
    const float tt[3] = {Tr[3] * sg, Tr[7] * sg, Tr[11] * sg};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) YR[i * 3 + j] = Ra[i * 3] * Tr[0 * 4 + j] + Ra[i * 3 + 1] * Tr[1 * 4 + j] + Ra[i * 3 + 2] * Tr[2 * 4 + j];
      Yt[i] = Ra[i * 3] * tt[0] + Ra[i * 3 + 1] * tt[1] + Ra[i * 3 + 2] * tt[2] + Ta[i];
    }
  };

  // ---- phase 1: image poses
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float R[9], t[3], gR[9], gt[3];
    load_RT(sP + n * 12, R, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gR[i * 3] = (float)a.gpose[n * 12 + i * 4]; gR[i * 3 + 1] = (float)a.gpose[n * 12 + i * 4 + 1];
      gR[i * 3 + 2] = (float)a.gpose[n * 12 + i * 4 + 2]; gt[i] = (float)a.gpose[n * 12 + i * 4 + 3];
    }
    if (a.tsw > 0.f) {
      float R2[9], t2[3];
      if (n + 1 < N) { load_RT(sP + (n + 1) * 12, R2, t2); relpose_grad(R, t, R2, t2, a.tw, a.tsw, gR, gt, nullptr, nullptr); }
      if (n > 0) { load_RT(sP + (n - 1) * 12, R2, t2); relpose_grad(R2, t2, R, t, a.tw, a.tsw, nullptr, nullptr, gR, gt); }
    }
    if (phaseB) {
      for (int k = a.edge_ptr[n]; k < a.edge_ptr[n + 1]; ++k) {
        const int e = a.edge_idx[k], g = e / gs;
        if (a.valid_traj[g] == 0.f) continue;
        float YR[9], Yt[3];
        traj_pose(e, g, YR, Yt);
        relpose_grad(YR, Yt, R, t, a.tw, 0.005f, nullptr, nullptr, gR, gt);
      }
    }
    float gq[4];
    quat_backward(sQ + n * 5, sQ[n * 5 + 4], gR, gq);
    float* prm = a.im_poses + n * 7;
    float newp[7];
#pragma unroll
    for (int i = 0; i < 4; ++i) newp[i] = adam_step(prm[i], gq[i], mP + n * 7 + i, vP + n * 7 + i, lr, bc1A, bc2A);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      newp[4 + i] = adam_step(prm[4 + i], gt[i] * expf(fabsf(prm[4 + i])), mP + n * 7 + 4 + i, vP + n * 7 + 4 + i, lr, bc1A, bc2A);
#pragma unroll
    for (int i = 0; i < 7; ++i) prm[i] = newp[i];
    float Rn[9], qn[4], in;
    quat_to_R(newp, Rn, qn, &in);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      a.poses_out[n * 12 + i * 4] = Rn[i * 3]; a.poses_out[n * 12 + i * 4 + 1] = Rn[i * 3 + 1];
      a.poses_out[n * 12 + i * 4 + 2] = Rn[i * 3 + 2]; a.poses_out[n * 12 + i * 4 + 3] = sexpm1(newp[4 + i]);
    }
  }
  // ---- phase 2: windows (sim(3)), depth scale/shift, trajectory alignment
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float dls = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) dls += (float)a.gS[g * 12 + i] * sW[g * 12 + i];
    sRed[g] = dls * sWq[g * 6 + 5];
  }
  __syncthreads();
  float sumD = 0.f;
  for (int g = 0; g < G; ++g) sumD += sRed[g];
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float sgm = sWq[g * 6 + 5];
    float gR[9], gt[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gR[i * 3] = sgm * (float)a.gS[g * 12 + i * 4]; gR[i * 3 + 1] = sgm * (float)a.gS[g * 12 + i * 4 + 1];
      gR[i * 3 + 2] = sgm * (float)a.gS[g * 12 + i * 4 + 2]; gt[i] = sgm * (float)a.gS[g * 12 + i * 4 + 3];
    }
    float gq[4];
    quat_backward(sWq + g * 6, sWq[g * 6 + 4], gR, gq);
    float* prm = a.pw_poses + g * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) prm[i] = adam_step(prm[i], gq[i], mW + g * 8 + i, vW + g * 8 + i, lr, bc1A, bc2A);
#pragma unroll
    for (int i = 0; i < 3; ++i) prm[4 + i] = adam_step(prm[4 + i], gt[i] * expf(fabsf(prm[4 + i])), mW + g * 8 + 4 + i, vW + g * 8 + 4 + i, lr, bc1A, bc2A);
    prm[7] = adam_step(prm[7], sRed[g] - sumD / (float)G, mW + g * 8 + 7, vW + g * 8 + 7, lr, bc1A, bc2A);
    if (phaseB) {
      a.s_depth[g] = adam_step(a.s_depth[g], (float)a.gst[g * 2], mS + g, vS + g, lr, bc1B, bc2B);
      a.t_depth[g] = adam_step(a.t_depth[g], (float)a.gst[g * 2 + 1], mT + g, vT + g, lr, bc1B, bc2B);
      float gRa[9], gTa[3], gsig = 0.f;
#pragma unroll
      for (int i = 0; i < 9; ++i) gRa[i] = 0.f;
      gTa[0] = gTa[1] = gTa[2] = 0.f;
      if (a.valid_traj[g] != 0.f) {
        float Ra[9], Ta[3];
        load_RT(sA + g * 12, Ra, Ta);
        const float sg = sAq[g * 6 + 5];
        for (int k = 0; k < gs; ++k) {
          const int e = g * gs + k, n = a.e_img[e];
          float YR[9], Yt[3], R[9], t[3], gYR[9], gYt[3];
          traj_pose(e, g, YR, Yt);
          load_RT(sP + n * 12, R, t);
#pragma unroll
          for (int i = 0; i < 9; ++i) gYR[i] = 0.f;
          gYt[0] = gYt[1] = gYt[2] = 0.f;
          relpose_grad(YR, Yt, R, t, a.tw, 0.005f, gYR, gYt, nullptr, nullptr);
          const float* Tr = a.traj + (long long)e * 16;
          const float tr[3] = {Tr[3], Tr[7], Tr[11]};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j)  // Y_R = Ra R_tr ; Y_t = Ra (sigma t_tr) + Ta
              gRa[i * 3 + j] += gYR[i * 3] * Tr[j * 4] + gYR[i * 3 + 1] * Tr[j * 4 + 1] + gYR[i * 3 + 2] * Tr[j * 4 + 2] + gYt[i] * sg * tr[j];
            gTa[i] += gYt[i];
            gsig += gYt[i] * (Ra[i * 3] * tr[0] + Ra[i * 3 + 1] * tr[1] + Ra[i * 3 + 2] * tr[2]);
          }
        }
        gsig *= sg;  // d sigma / d p7 = sigma
      }
      float gqa[4];
      quat_backward(sAq + g * 6, sAq[g * 6 + 4], gRa, gqa);
      float* pa = a.ta_poses + g * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) pa[i] = adam_step(pa[i], gqa[i], mA + g * 8 + i, vA + g * 8 + i, lr, bc1B, bc2B);
#pragma unroll
      for (int i = 0; i < 3; ++i) pa[4 + i] = adam_step(pa[4 + i], gTa[i] * expf(fabsf(pa[4 + i])), mA + g * 8 + 4 + i, vA + g * 8 + 4 + i, lr, bc1B, bc2B);
      pa[7] = adam_step(pa[7], gsig, mA + g * 8 + 7, vA + g * 8 + 7, lr, bc1B, bc2B);
    }
    a.st_out[g * 3] = a.s_depth[g];
    a.st_out[g * 3 + 1] = a.t_depth[g];
  }
  if (threadIdx.x == 0) {
    const float phi = a.im_focal[0];
    const float invf = expf(-phi / a.focal_break);
    const float gphi = (float)a.gscal[0] * invf * (-1.0f / a.focal_break);
    const float np_ = adam_step(phi, gphi, mF, vF, lr, bc1A, bc2A);
    a.im_focal[0] = np_;
    a.invf_out[0] = expf(-np_ / a.focal_break);
  }
  __syncthreads();
  // ---- refreshed sim(3) matrices (the scale normalisation couples all windows)
  float mean7n = 0.f;
  for (int g = 0; g < G; ++g) mean7n += a.pw_poses[g * 8 + 7];
  mean7n /= (float)G;
  const float normFn = expf(a.log_base_scale - mean7n);
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float R[9], qn[4], in;
    quat_to_R(a.pw_poses + g * 8, R, qn, &in);
    const float sgm = expf(a.pw_poses[g * 8 + 7]) * normFn;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      a.S_out[g * 12 + i * 4] = sgm * R[i * 3]; a.S_out[g * 12 + i * 4 + 1] = sgm * R[i * 3 + 1];
      a.S_out[g * 12 + i * 4 + 2] = sgm * R[i * 3 + 2]; a.S_out[g * 12 + i * 4 + 3] = sgm * sexpm1(a.pw_poses[g * 8 + 4 + i]);
    }
  }
}


}  // namespace g4
