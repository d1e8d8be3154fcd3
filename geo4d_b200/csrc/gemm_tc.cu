// Tap-GEMM: one persistent, warp-specialised tcgen05 kernel for every dense
// contraction of the Geo4D U-Net / VAE (linear, 1x1 conv, 3x3 conv, temporal
// (3,1,1) conv, batched matmul).  See include/geo4d_b200.h for the contract.
//
//   warp 0 lane 0 : TMA producer   (4-D activation box + 3-D weight box per k-step, SWIZZLE_128B)
//   warp 1 lane 0 : tcgen05.mma issuer (M=128, N=BLOCK_N, K=16 per instruction; fp32 accum in TMEM)
//   warps 2..9    : epilogue (tcgen05.ld -> bias / row-bias / SiLU / GEGLU / residual -> bf16|fp32 store);
//                   two warps per TMEM lane quadrant, additive terms staged in smem ahead of the MMAs
//
// The 3x3 / temporal taps are NOT im2col'ed: each tap is one more set of k-steps whose TMA box is
// shifted by (dx, dy); the TMA unit zero-fills the out-of-bounds halo, so the activation is read
// from HBM/L2 exactly as stored.  Two TMEM accumulator stages let the epilogue of tile i overlap
// the MMAs of tile i+1.
#include "common.cuh"
#include "geo4d_b200.h"

namespace g4 {

struct GemmArgs {
  int W, H, N;
  int box_w, box_h, box_n;
  int tiles_x, tiles_y, tiles_n;
  int n_tiles;     // along output columns
  int num_taps;
  int tap_dx[9], tap_dy[9];
  int kc_per_tap;  // K / 64
  int n_out;
  int b_batched;
  uint32_t a_box_bytes;
  void* out;
  long long ldc;
  int out_fp32;
  float alpha;
  const float* bias;
  const float* row_bias;
  long long row_bias_ld;
  int rows_per_bias;
  int act;
  const void* residual;
  long long ldr;
  unsigned long long* trace;  // debug: [grid][8] globaltimer stamps (geo4d_debug_gemm_trace), normally null
};

__device__ __forceinline__ void trace_stamp(const GemmArgs& a, int slot) {
  if (a.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(size_t)blockIdx.x * 8 + slot] = t;
  }
}

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = BLOCK_N * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 160 ? 5 : 6);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*bias*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(320, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmArgs args) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = bars + STAGES;          // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;      // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);  // [2][256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_stamp(args, 0);  // NB: the stamp itself is a global write, hence after the wait
  pdl_grid_sync();  // everything above overlaps the previous kernel's tail; global memory is touched only below
  if (threadIdx.x == 0) trace_stamp(args, 1);

  const int m_tiles = args.tiles_x * args.tiles_y * args.tiles_n;
  const int total_tiles = m_tiles * args.n_tiles;
  const int k_iters = args.num_taps * args.kc_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % args.n_tiles;
        const int mt = tile / args.n_tiles;
        const int x0 = (mt % args.tiles_x) * args.box_w;
        const int y0 = ((mt / args.tiles_x) % args.tiles_y) * args.box_h;
        const int n0 = (mt / (args.tiles_x * args.tiles_y)) * args.box_n;
        for (int tap = 0; tap < args.num_taps; ++tap) {
          const int dx = args.tap_dx[tap], dy = args.tap_dy[tap];
          const int bz = args.b_batched ? n0 : tap;
          for (int kc = 0; kc < args.kc_per_tap; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sB = sA + Cfg::A_BYTES;
            mbar_expect_tx(&full[stage], args.a_box_bytes + Cfg::B_BYTES);
            tma_load_4d(sA, &tmA, &full[stage], kc * 64, x0 + dx, y0 + dy, n0);
            tma_load_3d(sB, &tmB, &full[stage], kc * 64, nt * BLOCK_N, bz);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (it == 0 && ki == 0) trace_stamp(args, 2);
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_sw128_desc(sA + k * 32, 16, 1024);
            const uint64_t bd = make_sw128_desc(sB + k * 32, 16, 1024);
            umma_ss(d_tmem, ad, bd, idesc, (ki | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (it == 0) trace_stamp(args, 3);
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue: 8 warps; warps e and e+4 share a TMEM lane quadrant and split its 32-column chunks
    const int e = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = e >> 2;           // chunk parity handled by this warp
    const int et = threadIdx.x - 64;   // 0..255 among the epilogue threads
    const int r = q * 32 + lane;       // accumulator row handled by this thread
    const int rows_in_tile = args.box_w * args.box_h * args.box_n;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = tile % args.n_tiles;
      const int mt = tile / args.n_tiles;
      const int x0 = (mt % args.tiles_x) * args.box_w;
      const int y0 = ((mt / args.tiles_x) % args.tiles_y) * args.box_h;
      const int n0 = (mt / (args.tiles_x * args.tiles_y)) * args.box_n;
      const int x = x0 + (r % args.box_w);
      const int y = y0 + ((r / args.box_w) % args.box_h);
      const int n = n0 + r / (args.box_w * args.box_h);
      const bool row_ok = (r < rows_in_tile) && (x < args.W) && (y < args.H) && (n < args.N);
      const long long row = ((long long)n * args.H + y) * args.W + x;
      const int col_base = nt * BLOCK_N;

      // Stage the per-column additive terms (bias + the emb row of this tile) in shared memory while the
      // MMAs of this tile are still running: the epilogue then never waits on a global load for them.
      float* sb = s_bias + acc * 256;
      bool rb_in_smem = false;
      {
        long long rb_row = 0;
        if (args.row_bias) {
          const long long first = ((long long)n0 * args.H + y0) * args.W + x0;
          const int nl = min(n0 + args.box_n, args.N) - 1, yl = min(y0 + args.box_h, args.H) - 1,
                    xl = min(x0 + args.box_w, args.W) - 1;
          const long long last = ((long long)nl * args.H + yl) * args.W + xl;
          rb_row = first / args.rows_per_bias;
          rb_in_smem = (last / args.rows_per_bias) == rb_row;
        }
        if (et < BLOCK_N) {
          const int col = col_base + et;
          float bv = 0.f;
          if (col < args.n_out) {
            if (args.bias) bv = __ldg(args.bias + col);
            if (rb_in_smem) bv += __ldg(args.row_bias + rb_row * args.row_bias_ld + col);
          }
          sb[et] = bv;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (it == 0 && et == 0) trace_stamp(args, 4);
      const uint32_t t_row = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);

      if (args.act == G4_ACT_GEGLU) {
        // column blocks alternate [32 value | 32 gate]; stored width is n_out/2
#pragma unroll 1
        for (int c = half; c < BLOCK_N / 64; c += 2) {
          uint32_t v[32], g[32];
          tmem_ld32(t_row + c * 64, v);
          tmem_ld32(t_row + c * 64 + 32, g);
          tmem_ld_wait();
          const int col0 = col_base + c * 64;           // in B-row (interleaved) space
          const int ocol0 = col0 >> 1;                  // stored column
          if (row_ok && col0 < args.n_out) {
            float o[32];
            const float4* b4 = reinterpret_cast<const float4*>(sb + c * 64);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 ba = b4[j4], bg = b4[8 + j4];
              const float bav[4] = {ba.x, ba.y, ba.z, ba.w}, bgv[4] = {bg.x, bg.y, bg.z, bg.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int j = 4 * j4 + k;
                const float a = fmaf(__uint_as_float(v[j]), args.alpha, bav[k]);
                const float b = fmaf(__uint_as_float(g[j]), args.alpha, bgv[k]);
                o[j] = a * gelu_erf_f(b);
              }
            }
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(args.out) + row * args.ldc + ocol0;
            uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 w;
              w.x = pack_bf16x2(o[8 * j + 0], o[8 * j + 1]);
              w.y = pack_bf16x2(o[8 * j + 2], o[8 * j + 3]);
              w.z = pack_bf16x2(o[8 * j + 4], o[8 * j + 5]);
              w.w = pack_bf16x2(o[8 * j + 6], o[8 * j + 7]);
              o4[j] = w;
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = half; c < BLOCK_N / 32; c += 2) {
          const int col0 = col_base + c * 32;
          const bool active = row_ok && col0 < args.n_out;
          const int ncols = min(32, args.n_out - col0);
          const bool vec_ok = (ncols == 32);
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          // residual loads are issued before waiting on the TMEM load so both latencies overlap
          uint4 rr[4];
          const __nv_bfloat16* rp = nullptr;
          bool res_vec = false;
          if (active && args.residual) {
            rp = reinterpret_cast<const __nv_bfloat16*>(args.residual) + row * args.ldr + col0;
            res_vec = vec_ok && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0);
            if (res_vec) {
              const uint4* r4 = reinterpret_cast<const uint4*>(rp);
#pragma unroll
              for (int j = 0; j < 4; ++j) rr[j] = r4[j];
            }
          }
          tmem_ld_wait();
          if (active) {
            float o[32];
            const float4* b4 = reinterpret_cast<const float4*>(sb + c * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bb = b4[j4];
              o[4 * j4 + 0] = fmaf(__uint_as_float(v[4 * j4 + 0]), args.alpha, bb.x);
              o[4 * j4 + 1] = fmaf(__uint_as_float(v[4 * j4 + 1]), args.alpha, bb.y);
              o[4 * j4 + 2] = fmaf(__uint_as_float(v[4 * j4 + 2]), args.alpha, bb.z);
              o[4 * j4 + 3] = fmaf(__uint_as_float(v[4 * j4 + 3]), args.alpha, bb.w);
            }
            if (args.row_bias && !rb_in_smem) {  // tile straddles two embedding rows: per-thread loads
              const float* rb = args.row_bias + (row / args.rows_per_bias) * args.row_bias_ld + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) o[j] += __ldg(rb + j);
            }
            if (args.act == G4_ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = silu_f(o[j]);
            }
            if (args.residual) {
              if (res_vec) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float2 f;
                  f = unpack_bf16x2(rr[j].x); o[8 * j + 0] += f.x; o[8 * j + 1] += f.y;
                  f = unpack_bf16x2(rr[j].y); o[8 * j + 2] += f.x; o[8 * j + 3] += f.y;
                  f = unpack_bf16x2(rr[j].z); o[8 * j + 4] += f.x; o[8 * j + 5] += f.y;
                  f = unpack_bf16x2(rr[j].w); o[8 * j + 6] += f.x; o[8 * j + 7] += f.y;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < ncols) o[j] += __bfloat162float(rp[j]);
              }
            }
            if (args.out_fp32) {
              float* op = reinterpret_cast<float*>(args.out) + row * args.ldc + col0;
              if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
                float4* o4 = reinterpret_cast<float4*>(op);
#pragma unroll
                for (int j = 0; j < 8; ++j) o4[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < ncols) op[j] = o[j];
              }
            } else {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(args.out) + row * args.ldc + col0;
              if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
                uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 w;
                  w.x = pack_bf16x2(o[8 * j + 0], o[8 * j + 1]);
                  w.y = pack_bf16x2(o[8 * j + 2], o[8 * j + 3]);
                  w.z = pack_bf16x2(o[8 * j + 4], o[8 * j + 5]);
                  w.w = pack_bf16x2(o[8 * j + 6], o[8 * j + 7]);
                  o4[j] = w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < ncols) op[j] = __float2bfloat16(o[j]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (et == 0) trace_stamp(args, it == 0 ? 5 : 6);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_stamp(args, 7);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BLOCK_N>
static int launch_tap_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int num_sms,
                           cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tap_gemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("tap_gemm<%d>: cudaFuncSetAttribute(smem=%d): %s", BLOCK_N, Cfg::SMEM_BYTES,
                     cudaGetErrorString(e));
      return G4_ERR_CUDA;
    }
    attr_set = true;
  }
  const int total = a.tiles_x * a.tiles_y * a.tiles_n * a.n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  launch_pdl(tap_gemm_kernel<BLOCK_N>, dim3(grid), dim3(320), Cfg::SMEM_BYTES, stream, tmA, tmB, a);
  return check_launch("tap_gemm");
}

int device_sm_count();
static unsigned long long* g_gemm_trace = nullptr;

}  // namespace g4

extern "C" void geo4d_debug_gemm_trace(void* buf) { g4::g_gemm_trace = reinterpret_cast<unsigned long long*>(buf); }

using namespace g4;

extern "C" int geo4d_tap_gemm(const g4_gemm_desc* d, g4_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!d || !d->a || !d->b || !d->out) { set_last_error("tap_gemm: null pointer"); return G4_ERR_BAD_ARG; }
  if (d->K <= 0 || d->K % 64) { set_last_error("tap_gemm: K=%d must be a positive multiple of 64", d->K); return G4_ERR_BAD_ARG; }
  if (d->num_taps < 1 || d->num_taps > 9) { set_last_error("tap_gemm: num_taps=%d", d->num_taps); return G4_ERR_BAD_ARG; }
  const int rows = d->box_w * d->box_h * d->box_n;
  if (d->box_w < 1 || d->box_h < 1 || d->box_n < 1 || rows > 128 || d->box_w > 256) {
    set_last_error("tap_gemm: bad box %dx%dx%d", d->box_w, d->box_h, d->box_n); return G4_ERR_BAD_ARG;
  }
  if (d->b_batched && (d->box_n != 1 || d->num_taps != 1)) { set_last_error("tap_gemm: batched B needs box_n=1, 1 tap"); return G4_ERR_BAD_ARG; }
  if (d->n_out < 1) { set_last_error("tap_gemm: n_out=%d", d->n_out); return G4_ERR_BAD_ARG; }
  if (d->act == G4_ACT_GEGLU && (d->n_out % 64 || d->out_fp32 || d->residual || d->row_bias || (d->ldc % 8) ||
                                 (reinterpret_cast<uintptr_t>(d->out) & 15))) {
    set_last_error("tap_gemm: GEGLU needs n_out%%64==0, bf16 out, ldc%%8==0, 16B-aligned out, no residual/row_bias");
    return G4_ERR_BAD_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(d->a) & 15) || (reinterpret_cast<uintptr_t>(d->b) & 15) ||
      (d->a_stride_w % 8) || (d->a_stride_h % 8) || (d->a_stride_n % 8)) {
    set_last_error("tap_gemm: A/B must be 16-byte aligned with strides multiple of 8 elements"); return G4_ERR_BAD_ARG;
  }

  // pick the column tile with a small cost model: a k-step of a 128 x BN tile moves (16 KB + BN*128 B) through
  // shared memory twice (TMA write + MMA read) at ~128 B/clk, the epilogue costs ~BN*8 cycles, and the grid is
  // persistent over `sms` CTAs.  Small-M layers (the 10x16 / 5x8 levels) therefore get narrower tiles so that
  // more SMs have work; wide layers keep 256 / 160 (160 divides every U-Net width 320*k without padding).
  const int sms = device_sm_count();
  if (sms <= 0) return G4_ERR_CUDA;
  const int n = d->n_out;
  int bn = 0;
  {
    const long long m_tiles = (long long)((d->W + d->box_w - 1) / d->box_w) * ((d->H + d->box_h - 1) / d->box_h) *
                              ((d->N + d->box_n - 1) / d->box_n);
    const long long k_iters = (long long)d->num_taps * (d->K / 64);
    const int cands[5] = {256, 160, 128, 64, 32};
    double best = 0;
    for (int ci = 0; ci < 5; ++ci) {
      const int c = cands[ci];
      if (d->act == G4_ACT_GEGLU && (c == 160 || c == 32 || n % c)) continue;
      if (c > 32 && n <= c / 2 && c != 64) continue;                 // do not pad tiny outputs to wide tiles
      const long long n_tiles = (n + c - 1) / c;
      const long long waves = (m_tiles * n_tiles + sms - 1) / sms;
      const double cost = (double)waves * ((double)k_iters * (256.0 + 2.0 * c) + 400.0 + 8.0 * c);
      if (bn == 0 || cost < best) { bn = c; best = cost; }
    }
    if (bn == 0) bn = (d->act == G4_ACT_GEGLU) ? 64 : 32;
  }

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)d->a_stride_w * 2, (uint64_t)d->a_stride_h * 2, (uint64_t)d->a_stride_n * 2};
    uint32_t box[4] = {64, (uint32_t)d->box_w, (uint32_t)d->box_h, (uint32_t)d->box_n};
    int rc = make_tmap_bf16(&tmA, d->a, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t nb = d->b_batched ? (uint64_t)d->N : (uint64_t)d->num_taps;
    uint64_t dims[3] = {(uint64_t)d->K, (uint64_t)d->n_out, nb};
    uint64_t strides[2] = {(uint64_t)d->K * 2, (uint64_t)d->K * 2 * (uint64_t)d->n_out};
    uint32_t box[3] = {64, (uint32_t)bn, 1};
    int rc = make_tmap_bf16(&tmB, d->b, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  GemmArgs a;
  a.W = d->W; a.H = d->H; a.N = d->N;
  a.box_w = d->box_w; a.box_h = d->box_h; a.box_n = d->box_n;
  a.tiles_x = (d->W + d->box_w - 1) / d->box_w;
  a.tiles_y = (d->H + d->box_h - 1) / d->box_h;
  a.tiles_n = (d->N + d->box_n - 1) / d->box_n;
  a.n_tiles = (n + bn - 1) / bn;
  a.num_taps = d->num_taps;
  for (int i = 0; i < 9; ++i) { a.tap_dx[i] = d->tap_dx[i]; a.tap_dy[i] = d->tap_dy[i]; }
  a.kc_per_tap = d->K / 64;
  a.n_out = n;
  a.b_batched = d->b_batched;
  a.a_box_bytes = (uint32_t)rows * 128u;
  a.out = d->out; a.ldc = d->ldc; a.out_fp32 = d->out_fp32;
  a.alpha = d->alpha;
  a.bias = d->bias; a.row_bias = d->row_bias; a.row_bias_ld = d->row_bias_ld;
  a.rows_per_bias = d->rows_per_bias > 0 ? d->rows_per_bias : 1;
  a.act = d->act; a.residual = d->residual; a.ldr = d->ldr;
  a.trace = g_gemm_trace;

  switch (bn) {
    case 32: return launch_tap_gemm<32>(tmA, tmB, a, sms, stream);
    case 64: return launch_tap_gemm<64>(tmA, tmB, a, sms, stream);
    case 128: return launch_tap_gemm<128>(tmA, tmB, a, sms, stream);
    case 160: return launch_tap_gemm<160>(tmA, tmB, a, sms, stream);
    default: return launch_tap_gemm<256>(tmA, tmB, a, sms, stream);
  }
}
