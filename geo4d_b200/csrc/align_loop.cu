// The global-alignment optimisation loop (LightPointCloudGroupOptimizer.forward + backward + Adam for
// iterations [it0, it1); dust3r/cloud_opt/optimizer_group.py:440-525, base_opt_group.py:553-626) as ONE
// persistent cooperative kernel per phase -- and, across GPUs, ONE fused compute + exchange kernel.
//
// Per iteration
//   1. dense part: every CTA owns fixed (image, pixel-chunk) units.  Per pixel it back-projects the log-depth,
//      loops over the windows observing the image, forms the point-map L1 / inverse-depth L1 terms and their
//      gradient, applies Adam to the log-depth IN PLACE and accumulates the matrix-form gradients w.r.t. the
//      image pose, the focal, each window's sim(3) and depth scale/shift.  A unit's sums are reduced over the
//      CTA in a fixed tree and stored as one row of `part` -- no atomics, so the result is bit-reproducible;
//   2. grid barrier;
//   3. CTA 0 folds the unit rows in unit order (fp64).  Multi-GPU (images sharded over the ranks): it pushes this
//      rank's record {pose gradients of its images | per-window sums | focal / loss sums} straight into every
//      peer's receive buffer over NVLink (plain stores to peer-mapped memory, then one release-store of an
//      iteration flag per peer), waits for the peers' flags and adds the records in RANK order, so every rank
//      holds bit-identical totals.  It then runs the O(N + G) small-parameter step (align_small.cuh; replicated)
//      which also refreshes the pose / sim(3) / focal matrices of the next iteration;
//   4. grid barrier.
// 500 iterations cost two launches (phase A: [0, 150), phase B: [150, 500)) instead of 1000 kernel + 2000 memset
// nodes, and the exchange costs ~2 KB per rank and iteration with no NCCL call inside the loop.
#include "align_small.cuh"
#include "geo4d_b200.h"

namespace g4 {

constexpr int LOOP_THREADS = 256;
constexpr int LOOP_KMAX = 8;          // windows observing one image (AL_KMAX of align.cu)
constexpr int LOOP_MAX_RANKS = 16;
constexpr int PART_STRIDE = 15 + 14 * LOOP_KMAX + 1;   // floats per unit row: pose 12 | scal 3 | per edge slot 14

struct LoopArgs {
  float* logd; float* adam_m; float* adam_v;   // [N][HW]
  const float* pred;                           // [E][HW][3]
  const float* weight;                         // [E][HW]
  const float* invd;                           // [E][HW] or null
  const int* edge_ptr; const int* edge_idx;    // images -> incident (window, frame) edges (CSR)
  const float* scal;                           // [iters][8]: {-, cx, cy, lr, bias_corr1, bias_corr2_sqrt, 1/area, phaseB}
  float* poses; float* S; float* invf; float* st;    // matrices the dense part reads; rewritten by the small step
  double* gpose; double* gS; double* gscal; double* gst;   // totals of the last iteration (small-step inputs)
  float* part;                                 // [units][PART_STRIDE]
  unsigned int* bar;                           // {arrival count, generation}
  int N, G, HW, W, gs;
  int n_lo, n_hi;                              // images owned by this rank
  int chunks;                                  // pixel chunks per image
  int it0, it1;
  int world, rank;
  int img_lo[LOOP_MAX_RANKS + 1];              // image partition over the ranks
  int max_ne;                                  // most windows observing one image
  int rec_doubles;                             // doubles per (rank, parity) record slot
  double* peer_rec[LOOP_MAX_RANKS];            // every rank's receive buffer [2][world][rec_doubles] (peer-mapped)
  unsigned long long* peer_flag[LOOP_MAX_RANKS];   // every rank's flag array [world]
  unsigned long long flag_base;                // flags carry flag_base + iteration + 1 (monotonic across calls)
  unsigned long long* dbg;                     // optional [8] accumulated %globaltimer deltas of CTA 0 (ns)
  SmallArgs small;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Sense-free counting barrier over the (co-resident, cooperative launch) grid.  gen only ever grows.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int target = gen + 1;
    __threadfence();
    if (atomicAdd(&bar[0], 1u) == gridDim.x - 1) {
      bar[0] = 0u;
      __threadfence();
      atomicExch(&bar[1], target);
    } else {
      long long t0 = clock64();
      while ((int)(ld_acquire_gpu(&bar[1]) - target) < 0) {
        __nanosleep(40);   // a waiting CTA shares its SM with a working one (CTA 0's serial section): do not steal issue slots
        if (clock64() - t0 > G4_MBAR_TIMEOUT_CYCLES) {
          printf("g4: align_loop grid barrier timeout (block %d gen %u)\n", (int)blockIdx.x, target);
          __trap();
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
  gen += 1;
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

// One (image, chunk) unit with NE incident edges.  sm: [NE][16] {S(12), s, t, depth_valid, -} of the edges' windows.
template <int NE>
__device__ __forceinline__ void dense_unit(const LoopArgs& a, const int n, const int chunk, const float* __restrict__ sc,
                                           const float* __restrict__ sm, const int* __restrict__ s_edge,
                                           float* __restrict__ red, float* __restrict__ out_row) {
  const float invf = __ldcg(a.invf), cx = sc[1], cy = sc[2], lr = sc[3], bc1 = sc[4], bc2s = sc[5], invA = sc[6];
  const bool phaseB = sc[7] != 0.f && a.invd != nullptr;
  float R[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) R[i] = __ldcg(a.poses + n * 12 + i);
  float acc[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) acc[i] = 0.f;
  float accS[NE][14];
#pragma unroll
  for (int k = 0; k < NE; ++k)
#pragma unroll
    for (int i = 0; i < 14; ++i) accS[k][i] = 0.f;

  const int quads = a.HW >> 2;
  const int per = (quads + a.chunks - 1) / a.chunks;
  const int q0 = chunk * per, q1 = min(q0 + per, quads);
  const long long base = (long long)n * a.HW;
  for (int q = q0 + (int)threadIdx.x; q < q1; q += LOOP_THREADS) {
    const int p0 = q << 2;
    const float4 ld4 = ldcg4(a.logd + base + p0);
    const float4 m4 = ldcg4(a.adam_m + base + p0);
    const float4 v4 = ldcg4(a.adam_v + base + p0);
    const float ldv[4] = {ld4.x, ld4.y, ld4.z, ld4.w};
    const float mv[4] = {m4.x, m4.y, m4.z, m4.w};
    const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
    float d[4], xc[4], yc[4], Xw0[4], Xw1[4], Xw2[4], g0[4], g1[4], g2[4], ginv[4], inv[4], du[4], dv[4];
    int py = p0 / a.W, px = p0 - py * a.W;   // one integer division per quad; the other three pixels follow by increment
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      d[j] = __expf(ldv[j]);
      du[j] = (float)px - cx;
      dv[j] = (float)py - cy;
      if (++px == a.W) { px = 0; ++py; }
      xc[j] = d[j] * du[j] * invf;
      yc[j] = d[j] * dv[j] * invf;
      Xw0[j] = R[0] * xc[j] + R[1] * yc[j] + R[2] * d[j] + R[3];
      Xw1[j] = R[4] * xc[j] + R[5] * yc[j] + R[6] * d[j] + R[7];
      Xw2[j] = R[8] * xc[j] + R[9] * yc[j] + R[10] * d[j] + R[11];
      g0[j] = g1[j] = g2[j] = ginv[j] = 0.f;
      inv[j] = 1.0f / (d[j] + 1e-6f);
    }
#pragma unroll
    for (int k = 0; k < NE; ++k) {
      const long long ep = (long long)s_edge[k] * a.HW + p0;
      const float4 pa = __ldg(reinterpret_cast<const float4*>(a.pred + 3 * ep));
      const float4 pb = __ldg(reinterpret_cast<const float4*>(a.pred + 3 * ep + 4));
      const float4 pc = __ldg(reinterpret_cast<const float4*>(a.pred + 3 * ep + 8));
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.weight + ep));
      float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (phaseB) r4 = __ldg(reinterpret_cast<const float4*>(a.invd + ep));
      const float pr[12] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, pc.x, pc.y, pc.z, pc.w};
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
      const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
      const float* Sg = sm + k * 16;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p0x = pr[3 * j], p1x = pr[3 * j + 1], p2x = pr[3 * j + 2];
        const float Y0 = Sg[0] * p0x + Sg[1] * p1x + Sg[2] * p2x + Sg[3];
        const float Y1 = Sg[4] * p0x + Sg[5] * p1x + Sg[6] * p2x + Sg[7];
        const float Y2 = Sg[8] * p0x + Sg[9] * p1x + Sg[10] * p2x + Sg[11];
        const float r0 = Xw0[j] - Y0, r1 = Xw1[j] - Y1, r2 = Xw2[j] - Y2;
        const float nr = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
        const float w = fminf(wv[j], 10.0f);
        acc[13] += w * nr;
        const float c = nr > 0.f ? w * invA / nr : 0.f;   // torch norm backward is 0 at the origin
        const float qa = c * r0, qb = c * r1, qc = c * r2;
        g0[j] += qa; g1[j] += qb; g2[j] += qc;
        accS[k][0] -= qa * p0x; accS[k][1] -= qa * p1x; accS[k][2] -= qa * p2x; accS[k][3] -= qa;
        accS[k][4] -= qb * p0x; accS[k][5] -= qb * p1x; accS[k][6] -= qb * p2x; accS[k][7] -= qb;
        accS[k][8] -= qc * p0x; accS[k][9] -= qc * p1x; accS[k][10] -= qc * p2x; accS[k][11] -= qc;
        if (phaseB) {
          const float sg = Sg[12], tg = Sg[13], okg = Sg[14];
          const float rho = rv[j];
          const float mk = (rho > 0.05f && okg != 0.f) ? 1.f : 0.f;
          const float res = inv[j] - (sg * rho + tg);
          acc[14] += mk * fabsf(res);
          const float sgn = (res > 0.f) ? 1.f : ((res < 0.f) ? -1.f : 0.f);
          const float gl = sgn * mk * 2.0f * invA;
          ginv[j] += gl;
          accS[k][12] -= gl * rho;
          accS[k][13] -= gl;
        }
      }
    }
    float nl[4], nm[4], nv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float jx = R[0] * du[j] * invf + R[1] * dv[j] * invf + R[2];
      const float jy = R[4] * du[j] * invf + R[5] * dv[j] * invf + R[6];
      const float jz = R[8] * du[j] * invf + R[9] * dv[j] * invf + R[10];
      const float gd = g0[j] * jx + g1[j] * jy + g2[j] * jz - ginv[j] * inv[j] * inv[j];
      const float grad = gd * d[j];
      nm[j] = 0.9f * mv[j] + 0.1f * grad;                // torch.optim.Adam, betas (0.9, 0.9), eps 1e-8
      nv[j] = 0.9f * vv[j] + 0.1f * grad * grad;
      nl[j] = ldv[j] - (lr / bc1) * nm[j] / (sqrtf(nv[j]) / bc2s + 1e-8f);
      acc[0] += g0[j] * xc[j]; acc[1] += g0[j] * yc[j]; acc[2] += g0[j] * d[j];
      acc[3] += g1[j] * xc[j]; acc[4] += g1[j] * yc[j]; acc[5] += g1[j] * d[j];
      acc[6] += g2[j] * xc[j]; acc[7] += g2[j] * yc[j]; acc[8] += g2[j] * d[j];
      acc[9] += g0[j]; acc[10] += g1[j]; acc[11] += g2[j];
      acc[12] += d[j] * (g0[j] * (R[0] * du[j] + R[1] * dv[j]) + g1[j] * (R[4] * du[j] + R[5] * dv[j]) +
                         g2[j] * (R[8] * du[j] + R[9] * dv[j]));
    }
    stcg4(a.adam_m + base + p0, make_float4(nm[0], nm[1], nm[2], nm[3]));
    stcg4(a.adam_v + base + p0, make_float4(nv[0], nv[1], nv[2], nv[3]));
    stcg4(a.logd + base + p0, make_float4(nl[0], nl[1], nl[2], nl[3]));
  }

  // ---- fixed-tree reduction over the CTA: pose(12) | scal(3) | NE x 14
  constexpr int NV = 15 + 14 * NE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    float pose[12] = {acc[0], acc[1], acc[2], acc[9], acc[3], acc[4], acc[5], acc[10], acc[6], acc[7], acc[8], acc[11]};
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      float s = pose[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) red[warp * NV + i] = s;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float s = acc[12 + i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) red[warp * NV + 12 + i] = s;
    }
#pragma unroll
    for (int k = 0; k < NE; ++k)
#pragma unroll
      for (int i = 0; i < 14; ++i) {
        float s = accS[k][i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp * NV + 15 + k * 14 + i] = s;
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV; i += LOOP_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LOOP_THREADS / 32; ++w) s += red[w * NV + i];
    __stcg(out_row + i, s);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LOOP_THREADS, 2)
align_loop_kernel(const LoopArgs a) {
  extern __shared__ float lsh[];
  float* red = lsh;                                     // [16 warps][PART_STRIDE]
  float* sm = red + (LOOP_THREADS / 32) * PART_STRIDE;  // [LOOP_KMAX][16]
  float* small_sh = sm + LOOP_KMAX * 16;                // align_small_body scratch (CTA 0 only)
  double* img_sum = reinterpret_cast<double*>(small_sh + ((align_small_smem_floats(a.N, a.G) + 1) & ~(size_t)1));   // [n_loc][nvmax]
  __shared__ int s_edge[LOOP_KMAX];
  __shared__ double s_tot[3];
  const int n_loc = a.n_hi - a.n_lo;
  const int units = n_loc * a.chunks;
  unsigned int gen = ld_acquire_gpu(&a.bar[1]);   // identical in every CTA: the previous launch left it settled

  unsigned long long tk[7];
  auto stamp = [&](int i) {
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tk[i]));
  };
  for (int it = a.it0; it < a.it1; ++it) {
    const float* sc = a.scal + (long long)it * 8;
    stamp(0);
    // ------------------------------------------------------------------ 1. dense part
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int n = a.n_lo + u / a.chunks, chunk = u % a.chunks;
      const int e0 = a.edge_ptr[n], ne = min(a.edge_ptr[n + 1] - e0, LOOP_KMAX);
      if (threadIdx.x < ne) s_edge[threadIdx.x] = a.edge_idx[e0 + threadIdx.x];
      if (threadIdx.x < ne * 16) {
        const int k = threadIdx.x >> 4, i = threadIdx.x & 15;
        const int g = a.edge_idx[e0 + k] / a.gs;
        float v = 0.f;
        if (i < 12) v = __ldcg(a.S + g * 12 + i);
        else if (i < 15) v = __ldcg(a.st + g * 3 + (i - 12));
        sm[k * 16 + i] = v;
      }
      __syncthreads();
      float* row = a.part + (long long)u * PART_STRIDE;
      switch (ne) {
        case 1: dense_unit<1>(a, n, chunk, sc, sm, s_edge, red, row); break;
        case 2: dense_unit<2>(a, n, chunk, sc, sm, s_edge, red, row); break;
        case 3: dense_unit<3>(a, n, chunk, sc, sm, s_edge, red, row); break;
        case 4: dense_unit<4>(a, n, chunk, sc, sm, s_edge, red, row); break;
        case 5: dense_unit<5>(a, n, chunk, sc, sm, s_edge, red, row); break;
        case 6: dense_unit<6>(a, n, chunk, sc, sm, s_edge, red, row); break;
        case 7: dense_unit<7>(a, n, chunk, sc, sm, s_edge, red, row); break;
        case 8: dense_unit<8>(a, n, chunk, sc, sm, s_edge, red, row); break;
        default: break;   // an image no window observes contributes nothing
      }
    }
    // ------------------------------------------------------------------ 2.
    stamp(1);
    grid_barrier(a.bar, gen);
    stamp(2);
    // ------------------------------------------------------------------ 3. fold, exchange, small step
    if (blockIdx.x == 0) {
      const int N = a.N, G = a.G;
      // zero the totals (every entry is rewritten below; images of other ranks arrive through the exchange)
      for (int i = threadIdx.x; i < N * 12; i += LOOP_THREADS) a.gpose[i] = 0.0;
      for (int i = threadIdx.x; i < G * 12; i += LOOP_THREADS) a.gS[i] = 0.0;
      for (int i = threadIdx.x; i < G * 2; i += LOOP_THREADS) a.gst[i] = 0.0;
      if (threadIdx.x < 3) s_tot[threadIdx.x] = 0.0;
      __syncthreads();
      // (a) per-image sums: thread (local image, value) adds the image's chunk rows in chunk order.  All loads of a
      //     batch are issued before the first add, so the fold costs one or two L2 round trips, not `chunks` of them.
      const int nvmax = 15 + 14 * a.max_ne;
      for (int o = threadIdx.x; o < n_loc * nvmax; o += LOOP_THREADS) {
        const int nl = o / nvmax, i = o % nvmax;
        const float* src = a.part + (long long)nl * a.chunks * PART_STRIDE + i;
        double s = 0.0;
        for (int c = 0; c < a.chunks; c += 8) {   // (a partial last batch is predicated, not serialised)
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (c + j < a.chunks) ? __ldcg(src + (long long)(c + j) * PART_STRIDE) : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) s += (double)v[j];
        }
        img_sum[o] = s;
      }
      __syncthreads();
      for (int o = threadIdx.x; o < n_loc * 12; o += LOOP_THREADS)
        a.gpose[a.n_lo * 12 + o] = img_sum[(o / 12) * nvmax + (o % 12)];
      if (threadIdx.x < 3) {
        double s = 0.0;
        for (int nl = 0; nl < n_loc; ++nl) s += img_sum[nl * nvmax + 12 + threadIdx.x];
        s_tot[threadIdx.x] = s;
      }
      // (b) per-window sums: thread (g, i) walks the local images in order and picks the slot of window g
      for (int o = threadIdx.x; o < G * 14; o += LOOP_THREADS) {
        const int g = o / 14, i = o % 14;
        double s = 0.0;
        for (int nl = 0; nl < n_loc; ++nl) {
          const int n = a.n_lo + nl;
          const int e0 = a.edge_ptr[n], ne = min(a.edge_ptr[n + 1] - e0, LOOP_KMAX);
          for (int k = 0; k < ne; ++k)
            if (a.edge_idx[e0 + k] / a.gs == g) s += img_sum[nl * nvmax + 15 + k * 14 + i];
        }
        if (i < 12) a.gS[g * 12 + i] = s; else a.gst[g * 2 + (i - 12)] = s;
      }
      __syncthreads();
      stamp(3);
      if (a.world > 1) {
        // (c) exchange: record = {gpose of my images (max_loc*12) | gS (G*12) | gst (G*2) | scal (3)}
        const int par = it & 1;
        const int max_loc = (a.rec_doubles - G * 14 - 3) / 12;
        const unsigned long long flag = a.flag_base + (unsigned long long)it + 1ull;
        for (int o = threadIdx.x; o < a.rec_doubles; o += LOOP_THREADS) {
          double v = 0.0;
          if (o < max_loc * 12) { if (o < n_loc * 12) v = a.gpose[a.n_lo * 12 + o]; }
          else if (o < max_loc * 12 + G * 12) v = a.gS[o - max_loc * 12];
          else if (o < max_loc * 12 + G * 14) v = a.gst[o - max_loc * 12 - G * 12];
          else v = s_tot[o - max_loc * 12 - G * 14];
          for (int r = 0; r < a.world; ++r)   // NVLink stores into every peer's (and my own) receive buffer
            a.peer_rec[r][((size_t)par * a.world + a.rank) * a.rec_doubles + o] = v;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < a.world) st_release_sys(a.peer_flag[threadIdx.x] + a.rank, flag);
        if (threadIdx.x < a.world) {
          const unsigned long long* f = a.peer_flag[a.rank] + threadIdx.x;
          long long t0 = clock64();
          while (ld_acquire_sys(f) < flag) {
            if (clock64() - t0 > 8 * G4_MBAR_TIMEOUT_CYCLES) {
              printf("g4: align_loop rank %d: no record from rank %d for iteration %d\n", a.rank, (int)threadIdx.x, it);
              __trap();
            }
          }
        }
        __syncthreads();
        const double* mine = a.peer_rec[a.rank] + (size_t)par * a.world * a.rec_doubles;
        for (int r = 0; r < a.world; ++r) {
          const int lo = a.img_lo[r], cnt = a.img_lo[r + 1] - lo;
          const double* rec = mine + (size_t)r * a.rec_doubles;
          for (int o = threadIdx.x; o < cnt * 12; o += LOOP_THREADS) a.gpose[lo * 12 + o] = __ldcv(rec + o);
        }
        for (int o = threadIdx.x; o < G * 14 + 3; o += LOOP_THREADS) {
          double s = 0.0;
          for (int r = 0; r < a.world; ++r) s += __ldcv(mine + (size_t)r * a.rec_doubles + max_loc * 12 + o);   // rank order
          if (o < G * 12) a.gS[o] = s;
          else if (o < G * 14) a.gst[o - G * 12] = s;
          else a.gscal[o - G * 14] = s;
        }
      } else if (threadIdx.x < 3) {
        a.gscal[threadIdx.x] = s_tot[threadIdx.x];
      }
      __syncthreads();
      stamp(4);
      // (d) O(N + G) parameters + refreshed matrices (replicated on every rank from identical totals)
      align_small_body(a.small, it, small_sh);
      __syncthreads();
      stamp(5);
    }
    // ------------------------------------------------------------------ 4.
    grid_barrier(a.bar, gen);
    stamp(6);
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) {   // dense | barrier 1 | fold | exchange | small step | barrier 2
      for (int i = 0; i < 6; ++i) a.dbg[i] += tk[i + 1] - tk[i];
      a.dbg[6] += 1;
    }
  }
}

int device_sm_count();

}  // namespace g4

using namespace g4;

extern "C" size_t geo4d_align_loop_part_floats(int n_images_local, int chunks) {
  return (size_t)n_images_local * (size_t)chunks * PART_STRIDE;
}

extern "C" int geo4d_align_loop_record_doubles(int max_images_per_rank, int G) { return max_images_per_rank * 12 + G * 14 + 3; }

static size_t loop_smem_bytes(int N, int G, int n_loc, int max_ne) {
  const size_t floats = (size_t)(LOOP_THREADS / 32) * PART_STRIDE + LOOP_KMAX * 16 + ((align_small_smem_floats(N, G) + 1) & ~(size_t)1);
  return sizeof(float) * floats + sizeof(double) * (size_t)n_loc * (15 + 14 * max_ne);
}

// Chunks per image so that one rank's units fill the co-resident grid about once (fewest partial rows to fold).
extern "C" int geo4d_align_loop_chunks(int n_images_local, int HW) {
  const int sms = device_sm_count();
  if (sms <= 0 || n_images_local < 1) return 1;
  int c = 2 * sms / n_images_local;   // two CTAs per SM; floor: the units of a rank must not spill into a second wave
  const int quads = HW / 4;
  const int maxc = (quads + LOOP_THREADS - 1) / LOOP_THREADS;   // at least one quad per thread and chunk
  if (c > maxc) c = maxc;
  if (c < 1) c = 1;
  return c;
}

extern "C" int geo4d_align_loop(const g4_align_loop_desc* d, g4_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!d || !d->logd || !d->adam_m || !d->adam_v || !d->pred || !d->weight || !d->edge_ptr || !d->edge_idx || !d->scal ||
      !d->poses || !d->S || !d->invf || !d->st || !d->gpose || !d->gS || !d->gscal || !d->gst || !d->part || !d->bar ||
      !d->im_poses || !d->im_focal || !d->pw_poses || !d->s_depth || !d->t_depth || !d->ta_poses || !d->adam_small ||
      !d->traj || !d->e_img || !d->valid_traj) {
    set_last_error("align_loop: null pointer"); return G4_ERR_BAD_ARG;
  }
  if (d->N < 1 || d->N > 65535 || d->G < 1 || d->HW < 4 || (d->HW & 3) || d->chunks < 1 || d->it1 < d->it0 ||
      d->n_lo < 0 || d->n_hi > d->N || d->n_hi < d->n_lo) {
    set_last_error("align_loop: bad sizes (N=%d G=%d HW=%d chunks=%d it=[%d,%d) images=[%d,%d)); HW must be a multiple of 4",
                   d->N, d->G, d->HW, d->chunks, d->it0, d->it1, d->n_lo, d->n_hi);
    return G4_ERR_BAD_ARG;
  }
  if (d->max_edges_per_image > LOOP_KMAX) {
    set_last_error("align_loop: an image is observed by %d windows; at most %d supported", d->max_edges_per_image, LOOP_KMAX);
    return G4_ERR_UNSUPPORTED;
  }
  if (d->world < 1 || d->world > LOOP_MAX_RANKS || d->rank < 0 || d->rank >= d->world) {
    set_last_error("align_loop: world=%d rank=%d (at most %d ranks)", d->world, d->rank, LOOP_MAX_RANKS); return G4_ERR_BAD_ARG;
  }
  if (d->it1 == d->it0) return G4_OK;
  LoopArgs a;
  a.logd = d->logd; a.adam_m = d->adam_m; a.adam_v = d->adam_v; a.pred = d->pred; a.weight = d->weight; a.invd = d->invd;
  a.edge_ptr = d->edge_ptr; a.edge_idx = d->edge_idx; a.scal = d->scal;
  a.poses = d->poses; a.S = d->S; a.invf = d->invf; a.st = d->st;
  a.gpose = d->gpose; a.gS = d->gS; a.gscal = d->gscal; a.gst = d->gst;
  a.part = d->part; a.bar = d->bar;
  a.N = d->N; a.G = d->G; a.HW = d->HW; a.W = d->W; a.gs = d->group_size;
  a.n_lo = d->n_lo; a.n_hi = d->n_hi; a.chunks = d->chunks; a.it0 = d->it0; a.it1 = d->it1;
  a.world = d->world; a.rank = d->rank;
  a.rec_doubles = d->rec_doubles;
  a.flag_base = d->flag_base;
  a.dbg = reinterpret_cast<unsigned long long*>(d->debug_ns);
  for (int r = 0; r <= LOOP_MAX_RANKS; ++r) a.img_lo[r] = 0;
  for (int r = 0; r < LOOP_MAX_RANKS; ++r) { a.peer_rec[r] = nullptr; a.peer_flag[r] = nullptr; }
  if (d->world > 1) {
    int max_loc = 0;
    for (int r = 0; r <= d->world; ++r) a.img_lo[r] = d->img_lo[r];
    for (int r = 0; r < d->world; ++r) {
      if (!d->peer_rec[r] || !d->peer_flag[r]) { set_last_error("align_loop: peer buffer of rank %d is null", r); return G4_ERR_BAD_ARG; }
      a.peer_rec[r] = reinterpret_cast<double*>(d->peer_rec[r]);
      a.peer_flag[r] = reinterpret_cast<unsigned long long*>(d->peer_flag[r]);
      const int c = a.img_lo[r + 1] - a.img_lo[r];
      if (c < 0) { set_last_error("align_loop: image partition must be non-decreasing"); return G4_ERR_BAD_ARG; }
      if (c > max_loc) max_loc = c;
    }
    if (a.img_lo[0] != 0 || a.img_lo[d->world] != d->N || a.img_lo[d->rank] != d->n_lo || a.img_lo[d->rank + 1] != d->n_hi ||
        d->rec_doubles != max_loc * 12 + d->G * 14 + 3) {
      set_last_error("align_loop: inconsistent image partition / record size"); return G4_ERR_BAD_ARG;
    }
  } else {
    a.img_lo[1] = d->N;
    if (d->n_lo != 0 || d->n_hi != d->N) { set_last_error("align_loop: a single rank owns every image"); return G4_ERR_BAD_ARG; }
  }
  SmallArgs& s = a.small;
  s.im_poses = d->im_poses; s.im_focal = d->im_focal; s.pw_poses = d->pw_poses; s.s_depth = d->s_depth; s.t_depth = d->t_depth;
  s.ta_poses = d->ta_poses; s.adam = d->adam_small; s.gpose = d->gpose; s.gS = d->gS; s.gscal = d->gscal; s.gst = d->gst;
  s.traj = d->traj; s.e_img = d->e_img; s.edge_ptr = d->edge_ptr; s.edge_idx = d->edge_idx; s.valid_traj = d->valid_traj;
  s.scal = d->scal; s.it = nullptr; s.poses_out = d->poses; s.S_out = d->S; s.invf_out = d->invf; s.st_out = d->st;
  s.N = d->N; s.G = d->G; s.gs = d->group_size; s.start_b = d->start_b;
  s.tsw = d->temporal_smoothing_weight; s.tw = d->translation_weight; s.log_base_scale = logf(d->base_scale);
  s.focal_break = d->focal_break;

  a.max_ne = d->max_edges_per_image < 1 ? 1 : d->max_edges_per_image;
  const size_t smem = loop_smem_bytes(d->N, d->G, d->n_hi - d->n_lo, a.max_ne);
  if (smem > 200 * 1024) {
    set_last_error("align_loop: %d images (%d on this rank, up to %d windows each) / %d windows need %zu bytes of shared memory in one CTA",
                   d->N, d->n_hi - d->n_lo, a.max_ne, d->G, smem);
    return G4_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(align_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_last_error("align_loop: smem attr: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  const int sms = device_sm_count(); if (sms <= 0) return G4_ERR_CUDA;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, align_loop_kernel, LOOP_THREADS, smem);
  if (e != cudaSuccess || per_sm < 1) { set_last_error("align_loop: occupancy query failed"); (void)cudaGetLastError(); return G4_ERR_CUDA; }
  const int units = (d->n_hi - d->n_lo) * d->chunks;
  int grid = sms * per_sm;
  if (grid > units) grid = units;
  if (grid < 1) grid = 1;   // a rank that owns no image still takes part in the exchange and the small step
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(LOOP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, align_loop_kernel, a);
  if (e != cudaSuccess) { set_last_error("align_loop: launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return G4_ERR_CUDA; }
  return check_launch("align_loop");
}
