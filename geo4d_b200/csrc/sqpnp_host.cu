// Host-side SQPnP from reduced moments (fp64).  The per-pixel work of the reference's per-frame pose
// initialisation -- cv2.solvePnPRansac(flags=SOLVEPNP_SQPNP) inside fast_pnp, init_im_poses.py:824-865 -- is
// reduced on the GPU to 41 focal-independent moments per frame (geo4d_pnp_moments, align.cu).  What remains is a
// 9-unknown problem per (frame, tentative focal):  min r^T Omega r  over rotations r (rows of R in r), followed by
// t = P r.  This file solves it exactly like geo4d_b200/init_solvers.py:sqpnp_from_moments (Terzakis & Lourakis,
// "A consistently fast and globally optimal solution to the PnP problem", ECCV 2020): eigenvectors of Omega as
// starting points, sequential quadratic programming on SO(3), nearest-rotation projections, cheirality test.
// No CUDA here: it lives in the library so that the host side of the alignment does not pay ~1.3 ms of NumPy
// per solve (16 frames x 6 solves per window).
#include <math.h>
#include <string.h>

#include "geo4d_b200.h"

namespace {

// ---- small dense helpers (row-major)
inline void sym6(const double* v, double M[9]) {
  M[0] = v[0]; M[1] = v[1]; M[2] = v[2];
  M[3] = v[1]; M[4] = v[3]; M[5] = v[4];
  M[6] = v[2]; M[7] = v[4]; M[8] = v[5];
}

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 9); A is destroyed, V columns = vectors
void jacobi_eig(int n, double* A, double* V, double* w) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) off += A[p * n + q] * A[p * n + q];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p) {
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (fabs(apq) < 1e-300) continue;
        const double app = A[p * n + p], aqq = A[q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
    }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
  // ascending order (selection sort on the few entries)
  for (int i = 0; i < n; ++i) {
    int m = i;
    for (int j = i + 1; j < n; ++j)
      if (w[j] < w[m]) m = j;
    if (m != i) {
      const double tw = w[i]; w[i] = w[m]; w[m] = tw;
      for (int k = 0; k < n; ++k) { const double tv = V[k * n + i]; V[k * n + i] = V[k * n + m]; V[k * n + m] = tv; }
    }
  }
}

inline double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// nearest rotation to the 3x3 matrix E (row-major in e[9]): R = U diag(1, 1, det(U V^T)) V^T
void nearest_rotation(const double e[9], double r[9]) {
  double B[9], V[9], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B[i * 3 + j] = e[0 * 3 + i] * e[0 * 3 + j] + e[1 * 3 + i] * e[1 * 3 + j] + e[2 * 3 + i] * e[2 * 3 + j];
  jacobi_eig(3, B, V, w);                      // E^T E = V diag(w) V^T, ascending: column 2 = largest singular value
  // left vectors u_k = E v_k / sigma_k for the two largest singular values, the third by cross product
  double s2 = sqrt(fmax(w[2], 0.0)), s1 = sqrt(fmax(w[1], 0.0));
  double u2[3], u1[3], u0[3];
  for (int i = 0; i < 3; ++i) u2[i] = e[i * 3 + 0] * V[0 * 3 + 2] + e[i * 3 + 1] * V[1 * 3 + 2] + e[i * 3 + 2] * V[2 * 3 + 2];
  for (int i = 0; i < 3; ++i) u1[i] = e[i * 3 + 0] * V[0 * 3 + 1] + e[i * 3 + 1] * V[1 * 3 + 1] + e[i * 3 + 2] * V[2 * 3 + 1];
  if (s2 < 1e-300) { for (int i = 0; i < 9; ++i) r[i] = (i % 4 == 0) ? 1.0 : 0.0; return; }
  for (int i = 0; i < 3; ++i) u2[i] /= s2;
  if (s1 > 1e-12 * s2) {
    for (int i = 0; i < 3; ++i) u1[i] /= s1;
    double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];   // re-orthogonalise against u2
    for (int i = 0; i < 3; ++i) u1[i] -= d * u2[i];
  } else {                                                        // rank 1: any unit vector orthogonal to u2
    const int k = fabs(u2[0]) < fabs(u2[1]) ? (fabs(u2[0]) < fabs(u2[2]) ? 0 : 2) : (fabs(u2[1]) < fabs(u2[2]) ? 1 : 2);
    double a[3] = {0, 0, 0}; a[k] = 1.0;
    double d = a[0] * u2[0] + a[1] * u2[1] + a[2] * u2[2];
    for (int i = 0; i < 3; ++i) u1[i] = a[i] - d * u2[i];
  }
  double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  for (int i = 0; i < 3; ++i) u1[i] /= n1;
  // third left / right vectors complete right-handed frames, so U' V'^T with U' = [u0 u1 u2], V' = [v0 v1 v2] and
  // u0 = u1 x u2, v0 = v1 x v2 is the rotation closest to E (the det fix lands on the smallest singular value)
  u0[0] = u1[1] * u2[2] - u1[2] * u2[1]; u0[1] = u1[2] * u2[0] - u1[0] * u2[2]; u0[2] = u1[0] * u2[1] - u1[1] * u2[0];
  double v1[3] = {V[0 * 3 + 1], V[1 * 3 + 1], V[2 * 3 + 1]}, v2[3] = {V[0 * 3 + 2], V[1 * 3 + 2], V[2 * 3 + 2]}, v0[3];
  v0[0] = v1[1] * v2[2] - v1[2] * v2[1]; v0[1] = v1[2] * v2[0] - v1[0] * v2[2]; v0[2] = v1[0] * v2[1] - v1[1] * v2[0];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = u0[i] * v0[j] + u1[i] * v1[j] + u2[i] * v2[j];
}

// solve the n x n system K x = b in place (Gaussian elimination, partial pivoting); false if singular
bool solve_dense(int n, double* K, double* b) {
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(K[r * n + c]) > fabs(K[p * n + c])) p = r;
    if (fabs(K[p * n + c]) < 1e-300) return false;
    if (p != c) {
      for (int k = 0; k < n; ++k) { const double t = K[c * n + k]; K[c * n + k] = K[p * n + k]; K[p * n + k] = t; }
      const double t = b[c]; b[c] = b[p]; b[p] = t;
    }
    const double inv = 1.0 / K[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      const double f = K[r * n + c] * inv;
      if (f == 0.0) continue;
      for (int k = c; k < n; ++k) K[r * n + k] -= f * K[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double a = b[r];
    for (int k = r + 1; k < n; ++k) a -= K[r * n + k] * b[k];
    b[r] = a / K[r * n + r];
  }
  return true;
}

// sequential quadratic programming on  min r^T Omega r  s.t. r in SO(3): each step is the equality-constrained QP
// min (r+d)^T Omega (r+d) s.t. J d = -g, solved through its 15 x 15 KKT system
void sqp_refine(double r[9], const double Omega[81]) {
  for (int it = 0; it < 15; ++it) {
    const double* r1 = r; const double* r2 = r + 3; const double* r3 = r + 6;
    double g[6] = {r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2] - 1, r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2] - 1,
                   r3[0] * r3[0] + r3[1] * r3[1] + r3[2] * r3[2] - 1, r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2],
                   r1[0] * r3[0] + r1[1] * r3[1] + r1[2] * r3[2], r2[0] * r3[0] + r2[1] * r3[1] + r2[2] * r3[2]};
    double J[54];
    memset(J, 0, sizeof(J));
    for (int k = 0; k < 3; ++k) {
      J[0 * 9 + k] = 2 * r1[k]; J[1 * 9 + 3 + k] = 2 * r2[k]; J[2 * 9 + 6 + k] = 2 * r3[k];
      J[3 * 9 + k] = r2[k]; J[3 * 9 + 3 + k] = r1[k];
      J[4 * 9 + k] = r3[k]; J[4 * 9 + 6 + k] = r1[k];
      J[5 * 9 + 3 + k] = r3[k]; J[5 * 9 + 6 + k] = r2[k];
    }
    double K[225], rhs[15];
    memset(K, 0, sizeof(K));
    for (int i = 0; i < 9; ++i) {
      double a = 0.0;
      for (int j = 0; j < 9; ++j) { K[i * 15 + j] = Omega[i * 9 + j]; a += Omega[i * 9 + j] * r[j]; }
      rhs[i] = -a;
      for (int c = 0; c < 6; ++c) { K[i * 15 + 9 + c] = J[c * 9 + i]; K[(9 + c) * 15 + i] = J[c * 9 + i]; }
    }
    for (int c = 0; c < 6; ++c) rhs[9 + c] = -g[c];
    if (!solve_dense(15, K, rhs)) break;
    double dd = 0.0;
    for (int i = 0; i < 9; ++i) { r[i] += rhs[i]; dd += rhs[i] * rhs[i]; }
    if (dd < 1e-10) break;
  }
}

}  // namespace

// mom: the 41 moments of geo4d_pnp_moments for one frame; f: focal in pixels.  On success writes the world-to-camera
// rotation (row-major) and translation and returns 1; returns 0 when there is no valid solution (fewer than 4
// points, singular system, or no candidate in front of the camera).
extern "C" int geo4d_sqpnp_from_moments(const double* mom, double f, double* R_out, double* t_out) {
  if (!mom || !R_out || !t_out) return 0;
  const double n = mom[0];
  if (!(n >= 4) || !isfinite(f) || f <= 0) return 0;
  const double s1 = 1.0 / f, s2 = 1.0 / (f * f);
  double SQ[9] = {n, 0, -s1 * mom[1], 0, n, -s1 * mom[2], -s1 * mom[1], -s1 * mom[2], s2 * mom[3]};
  double Sm[3], Sxm[3], Sym[3], Srm[3];
  for (int i = 0; i < 3; ++i) { Sm[i] = mom[4 + i]; Sxm[i] = s1 * mom[7 + i]; Sym[i] = s1 * mom[10 + i]; Srm[i] = s2 * mom[13 + i]; }
  double QA[27];
  memset(QA, 0, sizeof(QA));
  for (int i = 0; i < 3; ++i) {
    QA[0 * 9 + i] = Sm[i]; QA[0 * 9 + 6 + i] = -Sxm[i];
    QA[1 * 9 + 3 + i] = Sm[i]; QA[1 * 9 + 6 + i] = -Sym[i];
    QA[2 * 9 + i] = -Sxm[i]; QA[2 * 9 + 3 + i] = -Sym[i]; QA[2 * 9 + 6 + i] = Srm[i];
  }
  double Mm[9], Mx[9], My[9], Mr[9];
  sym6(mom + 16, Mm); sym6(mom + 22, Mx); sym6(mom + 28, My); sym6(mom + 34, Mr);
  double AQA[81];
  memset(AQA, 0, sizeof(AQA));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      AQA[i * 9 + j] = Mm[i * 3 + j];
      AQA[(3 + i) * 9 + 3 + j] = Mm[i * 3 + j];
      AQA[(6 + i) * 9 + 6 + j] = s2 * Mr[i * 3 + j];
      AQA[i * 9 + 6 + j] = -s1 * Mx[i * 3 + j];
      AQA[(6 + i) * 9 + j] = -s1 * Mx[i * 3 + j];
      AQA[(3 + i) * 9 + 6 + j] = -s1 * My[i * 3 + j];
      AQA[(6 + i) * 9 + 3 + j] = -s1 * My[i * 3 + j];
    }
  // P = -SQ^{-1} QA  (three right-hand sides per column)
  double P[27];
  for (int c = 0; c < 9; ++c) {
    double K[9], b[3] = {QA[0 * 9 + c], QA[1 * 9 + c], QA[2 * 9 + c]};
    memcpy(K, SQ, sizeof(K));
    if (!solve_dense(3, K, b)) return 0;
    for (int i = 0; i < 3; ++i) P[i * 9 + c] = -b[i];
  }
  double Omega[81];
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      double a = AQA[i * 9 + j];
      for (int k = 0; k < 3; ++k) a += QA[k * 9 + i] * P[k * 9 + j];
      Omega[i * 9 + j] = a;
    }
  for (int i = 0; i < 9; ++i)
    for (int j = i + 1; j < 9; ++j) { const double a = 0.5 * (Omega[i * 9 + j] + Omega[j * 9 + i]); Omega[i * 9 + j] = a; Omega[j * 9 + i] = a; }
  double A[81], V[81], w[9];
  memcpy(A, Omega, sizeof(A));
  jacobi_eig(9, A, V, w);
  const double mean_pt[3] = {Sm[0] / n, Sm[1] / n, Sm[2] / n};
  bool have = false;
  double best_err = 0.0, best_R[9], best_t[3];
  for (int k = 0; k < 9; ++k) {
    if (k > 0 && have && !(best_err > 3.0 * w[k])) break;
    for (int sg = 0; sg < 2; ++sg) {
      const double sgn = sg == 0 ? 1.0 : -1.0;
      double e[9], r[9], rr[9];
      for (int i = 0; i < 9; ++i) e[i] = sgn * 1.7320508075688772 * V[i * 9 + k];
      nearest_rotation(e, r);
      sqp_refine(r, Omega);
      nearest_rotation(r, rr);
      double t[3];
      for (int i = 0; i < 3; ++i) {
        double a = 0.0;
        for (int j = 0; j < 9; ++j) a += P[i * 9 + j] * rr[j];
        t[i] = a;
      }
      if (!(rr[6] * mean_pt[0] + rr[7] * mean_pt[1] + rr[8] * mean_pt[2] + t[2] > 0)) continue;   // cheirality
      double err = 0.0;
      for (int i = 0; i < 9; ++i) {
        double a = 0.0;
        for (int j = 0; j < 9; ++j) a += Omega[i * 9 + j] * rr[j];
        err += rr[i] * a;
      }
      if (!have || err < best_err) {
        have = true; best_err = err;
        memcpy(best_R, rr, sizeof(best_R));
        memcpy(best_t, t, sizeof(best_t));
      }
    }
  }
  if (!have) return 0;
  memcpy(R_out, best_R, sizeof(best_R));
  memcpy(t_out, best_t, sizeof(best_t));
  return 1;
}

// Batch form: n independent (moments, focal) problems, solved on up to `threads` host threads (every problem is a
// few hundred microseconds of dense 9x9 / 15x15 algebra; the per-frame hypothesis sets of the PnP initialisation
// are solved together).  ok_out[i] = 1 on success.  HOST function, no CUDA.
#include <thread>
#include <vector>
extern "C" int geo4d_sqpnp_from_moments_batch(const double* mom, const double* f, int n, double* R_out, double* t_out,
                                              int* ok_out, int threads) {
  if (!mom || !f || !R_out || !t_out || !ok_out || n < 0) return 0;
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  auto work = [&](int w) {
    for (int i = w; i < n; i += threads)
      ok_out[i] = geo4d_sqpnp_from_moments(mom + (size_t)i * 41, f[i], R_out + (size_t)i * 9, t_out + (size_t)i * 3);
  };
  if (threads <= 1) { work(0); return 1; }
  std::vector<std::thread> pool;
  pool.reserve(threads - 1);
  for (int w = 1; w < threads; ++w) pool.emplace_back(work, w);
  work(0);
  for (auto& th : pool) th.join();
  return 1;
}
