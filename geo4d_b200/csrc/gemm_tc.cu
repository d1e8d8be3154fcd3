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
#include <cstring>

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
  int tma_store;              // 1: bf16 tile goes out through shared memory + TMA (tmC valid)
  int split_k;                // >1: K range split over `split_k` CTAs per output tile, raw fp32 partials to out + ks*split_stride
  long long split_stride;     // elements between the partial planes
  unsigned long long* trace;  // debug: [grid][16] globaltimer stamps (geo4d_debug_gemm_trace), normally null
};

__device__ __forceinline__ void trace_stamp(const GemmArgs& a, int slot) {
  if (a.trace) {
    unsigned long long t;
    if (slot < 8) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));   // comparable across SMs, ~0.25 us ticks
    else t = (unsigned long long)clock64();                                  // SM cycles, same-CTA deltas only
    a.trace[(size_t)blockIdx.x * 16 + slot] = t;
  }
}

template <int BLOCK_N, bool TWO>
struct GemmCfg {
  // TWO: the tile is 256 x BLOCK_N over a CTA pair; each CTA stages its own 128 A rows and HALF of the B rows
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = (TWO ? BLOCK_N / 2 : BLOCK_N) * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLAB_BYTES = 128 * 64;   // [128 rows][32 bf16], SWIZZLE_64B
  static constexpr int NSLAB = TWO ? ((BLOCK_N == 256) ? 1 : 2)
                                   : ((BLOCK_N == 256 || BLOCK_N == 128) ? 1 : 2);  // output slabs per epilogue half
  static constexpr int FIXED = 2 * NSLAB * SLAB_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*bias*/;
  static constexpr int FIT = (232448 - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = FIT > 8 ? 8 : FIT;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED;
  static_assert(STAGES >= 4 && SMEM_BYTES <= 232448, "tap_gemm: shared memory budget");
};

template <int BLOCK_N, bool TWO>
__global__ void __launch_bounds__(320, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GemmArgs args) {
  using Cfg = GemmCfg<BLOCK_N, TWO>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // align by OFFSET (pointer arithmetic on the __shared__ array keeps the address space: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slabs = smem + STAGES * Cfg::STAGE_BYTES;   // [2 halves][NSLAB][SLAB_BYTES], 1024-aligned
  uint8_t* tail = slabs + 2 * Cfg::NSLAB * Cfg::SLAB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = bars + STAGES;          // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;      // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* s_bias = reinterpret_cast<float*>(tail + 256);  // [2][256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = TWO ? (int)cluster_ctarank() : 0;           // 0 = leader of the pair
  const int unit = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // scheduling unit: CTA or CTA pair
  const int n_units = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (args.tma_store) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], TWO ? 16 : 8);   // the leader's barrier collects both CTAs' epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (TWO) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncwarp();
  if (TWO) cluster_sync_all();   // the partner's barriers must exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_stamp(args, 0);  // NB: the stamp itself is a global write, hence after the wait
  pdl_grid_sync();  // everything above overlaps the previous kernel's tail; global memory is touched only below
  if (threadIdx.x == 0) trace_stamp(args, 1);

  const int m_tiles = args.tiles_x * args.tiles_y * args.tiles_n;
  // a pair takes two consecutive row boxes (2*mt2 + rank); an odd tail box is fully out of bounds for the
  // partner: its loads are zero-filled and its stores clipped by the TMA unit
  const int out_tiles = (TWO ? (m_tiles + 1) / 2 : m_tiles) * args.n_tiles;
  const int total_tiles = out_tiles * args.split_k;   // tile = out_tile * split_k + ks: the parts of one tile run side by side
  const int k_iters = args.num_taps * args.kc_per_tap;
  const int k_per = (k_iters + args.split_k - 1) / args.split_k;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        const int ot = tile / args.split_k, ks = tile - ot * args.split_k;
        const int nt = ot % args.n_tiles;
        const int mt = TWO ? 2 * (ot / args.n_tiles) + rank : ot / args.n_tiles;
        const int x0 = (mt % args.tiles_x) * args.box_w;
        const int y0 = ((mt / args.tiles_x) % args.tiles_y) * args.box_h;
        const int n0 = (mt / (args.tiles_x * args.tiles_y)) * args.box_n;
        const int k0 = ks * k_per, k1 = min(k0 + k_per, k_iters);
        int tap = k0 / args.kc_per_tap, kc = k0 - tap * args.kc_per_tap;
        for (int ki = k0; ki < k1; ++ki) {
          const int dx = args.tap_dx[tap], dy = args.tap_dy[tap];
          const int bz = args.b_batched ? n0 : tap;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          if (TWO) {
            // both CTAs' bytes are counted on the leader's barrier; only the leader arms it
            if (rank == 0) mbar_expect_tx(&full[stage], 2u * (args.a_box_bytes + Cfg::B_BYTES));
            const uint32_t lb = leader_bar_addr(&full[stage]);
            tma_load_4d_2sm(sA, &tmA, lb, kc * 64, x0 + dx, y0 + dy, n0);
            tma_load_3d_2sm(sB, &tmB, lb, kc * 64, nt * BLOCK_N + rank * (BLOCK_N / 2), bz);
          } else {
            mbar_expect_tx(&full[stage], args.a_box_bytes + Cfg::B_BYTES);
            tma_load_4d(sA, &tmA, &full[stage], kc * 64, x0 + dx, y0 + dy, n0);
            tma_load_3d(sB, &tmB, &full[stage], kc * 64, nt * BLOCK_N, bz);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++kc == args.kc_per_tap) { kc = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {   // the leader issues the MMAs of the pair
      constexpr uint32_t idesc = make_idesc_bf16(TWO ? 256 : 128, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = unit; tile < total_tiles; tile += n_units, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const int ks_m = tile % args.split_k;
        const int kn = min(k_per, k_iters - ks_m * k_per);
        for (int ki = 0; ki < kn; ++ki) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (it == 0 && ki == 0) trace_stamp(args, 2);
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_sw128_desc(sA + k * 32, 16, 1024);
            const uint64_t bd = make_sw128_desc(sB + k * 32, 16, 1024);
            if (TWO) umma_ss_2sm(d_tmem, ad, bd, idesc, (ki | k) != 0 ? 1u : 0u);
            else umma_ss(d_tmem, ad, bd, idesc, (ki | k) != 0 ? 1u : 0u);
          }
          if (TWO) umma_commit_2sm(&empty[stage]); else umma_commit(&empty[stage]);   // frees the stage in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (TWO) umma_commit_2sm(&tfull[acc]); else umma_commit(&tfull[acc]);
        if (it == 0) trace_stamp(args, 3);
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue: 8 warps; warps e and e+4 share a TMEM lane quadrant and split its 32-column chunks.
    // bf16 results go through shared memory: each 4-warp half owns a small ring of [128 rows x 32 cols]
    // slabs (64-byte rows, SWIZZLE_64B so that the per-row 16-byte writes are bank-conflict free); a slab is
    // written by the 128 threads of the half and drained by ONE TMA store, which also clips the tile against
    // the tensor edges.  Direct per-thread stores remain for fp32 outputs and unaligned destinations.
    const int e = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = e >> 2;           // chunk parity handled by this warp
    const int et = threadIdx.x - 64;   // 0..255 among the epilogue threads
    const int r = q * 32 + lane;       // accumulator row handled by this thread
    const int rows_in_tile = args.box_w * args.box_h * args.box_n;
    const bool issuer = ((e & 3) == 0) && lane == 0;
    const int bar_id = 2 + half;
    uint8_t* my_slabs = slabs + half * (Cfg::NSLAB * Cfg::SLAB_BYTES);
    const uint32_t sw = (uint32_t)(r >> 1) & 3u;   // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,9)
    uint32_t kk = 0;                   // slabs this half has filled so far (ring position)
    const int n_chunks = (args.act == G4_ACT_GEGLU) ? BLOCK_N / 64 : BLOCK_N / 32;
    const int acc_cols = (args.act == G4_ACT_GEGLU) ? 64 : 32;   // accumulator columns per chunk
    int it = 0;
    for (int tile = unit; tile < total_tiles; tile += n_units, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int ot = tile / args.split_k, ks = tile - ot * args.split_k;
      const int nt = ot % args.n_tiles;
      const int mt = TWO ? 2 * (ot / args.n_tiles) + rank : ot / args.n_tiles;
      const int x0 = (mt % args.tiles_x) * args.box_w;
      const int y0 = ((mt / args.tiles_x) % args.tiles_y) * args.box_h;
      const int n0 = (mt / (args.tiles_x * args.tiles_y)) * args.box_n;
      const int x = x0 + (r % args.box_w);
      const int y = y0 + ((r / args.box_w) % args.box_h);
      const int n = n0 + r / (args.box_w * args.box_h);
      const bool row_ok = (r < rows_in_tile) && (x < args.W) && (y < args.H) && (n < args.N);
      const long long row = ((long long)n * args.H + y) * args.W + x;
      const int col_base = nt * BLOCK_N;
      // The two halves run decoupled (no block-wide barrier): half h owns the chunks of parity (h + it) & 1, so
      // an odd chunk count (BLOCK_N = 160) balances over two tiles.  c = par, par+2, ... inside n_out.
      const int par = (half + it) & 1;
      int my_chunks = 0;
      for (int c = par; c < n_chunks && col_base + c * acc_cols < args.n_out; c += 2) ++my_chunks;

      // Stage the per-column additive terms (bias + the emb row of this tile) of this half's chunks in shared
      // memory while the MMAs of this tile are still running: the epilogue then never waits on a global load.
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
        const int ht = (e & 3) * 32 + lane;              // 0..127 within the half
        if (ht < my_chunks * acc_cols) {
          const int lc = (par + 2 * (ht / acc_cols)) * acc_cols + ht % acc_cols;   // column inside the tile
          const int col = col_base + lc;
          float bv = 0.f;
          if (col < args.n_out) {
            if (args.bias) bv = __ldg(args.bias + col);
            if (rb_in_smem) bv += __ldg(args.row_bias + rb_row * args.row_bias_ld + col);
          }
          sb[lc] = bv;
        }
        named_bar_sync(bar_id, 128);
      }

      // residual of the first chunk: requested before the accumulator is ready, so its latency hides
      // behind the MMAs of this tile; each later chunk's residual is requested one chunk ahead.
      uint32_t rr[16];
      const __nv_bfloat16* rrow = nullptr;
      bool res_vec = false;
      if (args.residual && row_ok && my_chunks > 0) {
        rrow = reinterpret_cast<const __nv_bfloat16*>(args.residual) + row * args.ldr + col_base;
        res_vec = ((reinterpret_cast<uintptr_t>(rrow) & 31) == 0) && ((args.ldr & 15) == 0);
        if (res_vec && col_base + par * 32 + 32 <= args.n_out) {
          uint32_t(&lo)[8] = *reinterpret_cast<uint32_t(*)[8]>(&rr[0]);
          uint32_t(&hi)[8] = *reinterpret_cast<uint32_t(*)[8]>(&rr[8]);
          ldg256(rrow + par * 32, lo);
          ldg256(rrow + par * 32 + 16, hi);
        }
      }

      if (my_chunks == 0) {   // (narrow tiles: BLOCK_N == 32 has a single chunk) nothing to read from TMEM
        if (lane == 0) { if (TWO) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
        continue;
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (it == 0 && et == 0) { trace_stamp(args, 4); trace_stamp(args, 13); }
      const uint32_t t_row = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);

#pragma unroll 1
      for (int ci = 0; ci < my_chunks; ++ci) {
        const int c = par + 2 * ci;
        const int col0 = col_base + c * acc_cols;              // first accumulator column (B-row space)
        const int scol0 = (args.act == G4_ACT_GEGLU) ? (col0 >> 1) : col0;   // first stored column
        const int n_store = (args.act == G4_ACT_GEGLU) ? (args.n_out >> 1) : args.n_out;
        const int ncols = min(32, n_store - scol0);
        const bool vec_ok = (ncols == 32);
        float o[32];
        if (args.act == G4_ACT_GEGLU) {
          // column blocks alternate [32 value | 32 gate]; stored width is n_out/2
          uint32_t v[32], g[32];
          tmem_ld32(t_row + c * 64, v);
          tmem_ld32(t_row + c * 64 + 32, g);
          tmem_ld_wait();
          if (ci == my_chunks - 1) {   // accumulator fully read by this warp: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (TWO) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
          }
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
        } else {
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          if (it == 0 && ci == 0 && et == 0) trace_stamp(args, 8);
          if (ci == my_chunks - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (TWO) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
          }
          const float4* b4 = reinterpret_cast<const float4*>(sb + c * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bb = b4[j4];
            o[4 * j4 + 0] = fmaf(__uint_as_float(v[4 * j4 + 0]), args.alpha, bb.x);
            o[4 * j4 + 1] = fmaf(__uint_as_float(v[4 * j4 + 1]), args.alpha, bb.y);
            o[4 * j4 + 2] = fmaf(__uint_as_float(v[4 * j4 + 2]), args.alpha, bb.z);
            o[4 * j4 + 3] = fmaf(__uint_as_float(v[4 * j4 + 3]), args.alpha, bb.w);
          }
          if (args.row_bias && !rb_in_smem && row_ok) {  // tile straddles two embedding rows: per-thread loads
            const float* rb = args.row_bias + (row / args.rows_per_bias) * args.row_bias_ld + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) o[j] += __ldg(rb + j);
          }
          if (args.act == G4_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = silu_f(o[j]);
          } else if (args.act == G4_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = gelu_erf_f(o[j]);
          }
          if (rrow) {
            if (res_vec && vec_ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 f = unpack_bf16x2(rr[j]);
                o[2 * j] += f.x;
                o[2 * j + 1] += f.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) o[j] += __bfloat162float(rrow[c * 32 + j]);
            }
          }
        }

        if (args.tma_store) {
          uint4 w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            w[j].x = pack_bf16x2(o[8 * j + 0], o[8 * j + 1]);
            w[j].y = pack_bf16x2(o[8 * j + 2], o[8 * j + 3]);
            w[j].z = pack_bf16x2(o[8 * j + 4], o[8 * j + 5]);
            w[j].w = pack_bf16x2(o[8 * j + 6], o[8 * j + 7]);
          }
          if (it == 0 && ci == 0 && et == 0) trace_stamp(args, 9);
          {
            uint8_t* slab = my_slabs + (kk % Cfg::NSLAB) * Cfg::SLAB_BYTES;
            uint4* rowp = reinterpret_cast<uint4*>(slab + r * 64);
            if (Cfg::NSLAB >= 2) {
              // Two-slab ring, ONE barrier per chunk: the barrier of chunk k-1 already told everybody that this
              // slab is free, because the issuer waits for all earlier stores (<= k-2 at that point) to have
              // drained their slabs before it joins a barrier.
#pragma unroll
              for (int j = 0; j < 4; ++j) rowp[(uint32_t)j ^ sw] = w[j];
              fence_proxy_async_smem();
              if (issuer) bulk_wait_read<0>();
              named_bar_sync(bar_id, 128);
            } else {
              if (issuer) bulk_wait_read<0>();   // the store that last used this slab has drained it
              named_bar_sync(bar_id, 128);
#pragma unroll
              for (int j = 0; j < 4; ++j) rowp[(uint32_t)j ^ sw] = w[j];
              fence_proxy_async_smem();
              named_bar_sync(bar_id, 128);
            }
            if (it == 0 && ci == 0 && et == 0) trace_stamp(args, 11);
            if (issuer) {
              tma_store_4d(&tmC, slab, scol0, x0, y0, n0);
              bulk_commit();
            }
            if (it == 0 && ci == 0 && et == 0) trace_stamp(args, 12);
          }
          ++kk;
        } else if (row_ok) {
          if (args.out_fp32) {
            float* op = reinterpret_cast<float*>(args.out) + (long long)ks * args.split_stride + row * args.ldc + scol0;
            if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 31) == 0)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t pk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) pk[k] = __float_as_uint(o[8 * j + k]);
                stg256(op + 8 * j, pk);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncols) op[j] = o[j];
            }
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(args.out) + row * args.ldc + scol0;
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
        // next chunk's residual: requested only now, AFTER fence.proxy.async above -- that fence waits for the
        // thread's outstanding global loads, so an earlier request would put the L2 latency on the critical path
        if (rrow && ci + 1 < my_chunks && res_vec && args.act != G4_ACT_GEGLU &&
            col_base + (c + 2) * 32 + 32 <= args.n_out) {
          uint32_t(&lo)[8] = *reinterpret_cast<uint32_t(*)[8]>(&rr[0]);
          uint32_t(&hi)[8] = *reinterpret_cast<uint32_t(*)[8]>(&rr[8]);
          ldg256(rrow + (c + 2) * 32, lo);
          ldg256(rrow + (c + 2) * 32 + 16, hi);
        }
      }
      if (et == 0) trace_stamp(args, it == 0 ? 5 : 6);
    }
    if (issuer) bulk_wait_all();   // shared memory must outlive the last TMA store
  }

  tc_fence_before();
  __syncwarp();
  if (TWO) cluster_sync_all();   // neither CTA may retire while its partner can still touch its barriers / TMEM
  else __syncthreads();
  if (threadIdx.x == 0) trace_stamp(args, 7);
  if (warp == 1) {
    tc_fence_after();
    if (TWO) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}


// Second pass of a split-K tap-GEMM: out = epilogue(sum over the `split` fp32 partial planes, in plane order).
// Same epilogue terms as the fused path (alpha, bias, per-frame row bias, SiLU, residual); 8 columns per thread.
struct ReduceArgs {
  const float* ws; long long plane; int split;
  long long M; int n_out;
  void* out; long long ldc; int out_fp32;
  float alpha; const float* bias; const float* row_bias; long long row_bias_ld; int rows_per_bias; int act;
  const void* residual; long long ldr;
};

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const ReduceArgs a) {
  const int groups = (a.n_out + 7) >> 3;
  const long long total = a.M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / groups;
    const int c0 = (int)(i - row * groups) << 3;
    const int nc = min(8, a.n_out - c0);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* src = a.ws + row * a.n_out + c0;
    const bool vec = (nc == 8) && ((a.n_out & 3) == 0);
    for (int s = 0; s < a.split; ++s) {
      const float* p = src + (long long)s * a.plane;
      if (vec) {
        const float4 u = __ldcg(reinterpret_cast<const float4*>(p)), v = __ldcg(reinterpret_cast<const float4*>(p + 4));
        acc[0] += u.x; acc[1] += u.y; acc[2] += u.z; acc[3] += u.w;
        acc[4] += v.x; acc[5] += v.y; acc[6] += v.z; acc[7] += v.w;
      } else {
        for (int j = 0; j < nc; ++j) acc[j] += __ldcg(p + j);
      }
    }
    const float* rb = a.row_bias ? a.row_bias + (row / a.rows_per_bias) * a.row_bias_ld + c0 : nullptr;
    const __nv_bfloat16* rr = a.residual ? reinterpret_cast<const __nv_bfloat16*>(a.residual) + row * a.ldr + c0 : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < nc) {
        float o = fmaf(acc[j], a.alpha, a.bias ? __ldg(a.bias + c0 + j) : 0.f);
        if (rb) o += __ldg(rb + j);
        if (a.act == G4_ACT_SILU) o = silu_f(o);
        else if (a.act == G4_ACT_GELU) o = gelu_erf_f(o);
        if (rr) o += __bfloat162float(rr[j]);
        acc[j] = o;
      }
    }
    if (a.out_fp32) {
      float* op = reinterpret_cast<float*>(a.out) + row * a.ldc + c0;
      for (int j = 0; j < nc; ++j) op[j] = acc[j];
    } else {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + row * a.ldc + c0;
      if (nc == 8 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
        uint4 w;
        w.x = pack_bf16x2(acc[0], acc[1]); w.y = pack_bf16x2(acc[2], acc[3]);
        w.z = pack_bf16x2(acc[4], acc[5]); w.w = pack_bf16x2(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(op) = w;
      } else {
        for (int j = 0; j < nc; ++j) op[j] = __float2bfloat16(acc[j]);
      }
    }
  }
}

template <int BLOCK_N, bool TWO>
static int launch_tap_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                           const GemmArgs& a, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, TWO>;
  static unsigned long long attr_mask = 0;   // one latch per template instance, one bit per device
  if (once_per_device(&attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(tap_gemm_kernel<BLOCK_N, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      unlatch_device(&attr_mask);
      set_last_error("tap_gemm<%d,%d>: cudaFuncSetAttribute(smem=%d): %s", BLOCK_N, (int)TWO, Cfg::SMEM_BYTES,
                     cudaGetErrorString(e));
      return G4_ERR_CUDA;
    }
  }
  const int m_tiles = a.tiles_x * a.tiles_y * a.tiles_n;
  int grid;
  if (TWO) {
    const int total = ((m_tiles + 1) / 2) * a.n_tiles, pairs = num_sms / 2;
    grid = 2 * (total < pairs ? total : pairs);
  } else {
    const int total = m_tiles * a.n_tiles * a.split_k;
    grid = total < num_sms ? total : num_sms;
  }
  cudaError_t e = launch_ex(tap_gemm_kernel<BLOCK_N, TWO>, dim3(grid), dim3(320), Cfg::SMEM_BYTES, stream, TWO ? 2 : 1,
                            tmA, tmB, tmC, a);
  if (e != cudaSuccess) {
    set_last_error("tap_gemm<%d,%d>: launch: %s", BLOCK_N, (int)TWO, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return G4_ERR_CUDA;
  }
  return check_launch("tap_gemm");
}

int device_sm_count();
static unsigned long long* g_gemm_trace = nullptr;
static bool g_no_tma_store = false;   // debug switch: force the direct-store epilogue
static int g_pair_mode = -1;          // debug switch: -1 cost model, 0 never pair CTAs, 1 always (when legal)

}  // namespace g4

extern "C" void geo4d_debug_gemm_trace(void* buf) { g4::g_gemm_trace = reinterpret_cast<unsigned long long*>(buf); }
extern "C" void geo4d_debug_gemm_direct_store(int on) { g4::g_no_tma_store = on != 0; }
extern "C" void geo4d_debug_gemm_pair_mode(int mode) { g4::g_pair_mode = mode; }

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

  // Pick the column tile and 1-CTA vs CTA-pair with a small cost model.  A k-step of a CTA moves
  // (16 KB + its share of the B tile) from L2 into shared memory; the L2 -> SM path sustains ~45 B/clk per SM
  // (measured: profiles/), the MMA needs 2*BLOCK_N clk, so the k-step costs the larger of the two.  The epilogue
  // of a tile (~300 + 10*BLOCK_N clk) overlaps the next tile's main loop, so a CTA that runs `waves` tiles
  // takes main + (waves-1)*max(main, epi) + epi.  The grid is persistent over `sms` CTAs (or sms/2 pairs): a
  // pair halves the B bytes per SM but also the number of schedulable units, so epilogue-bound (small K) and
  // small-M layers (the 10x16 / 5x8 levels) keep single CTAs.
  const int sms = device_sm_count();
  if (sms <= 0) return G4_ERR_CUDA;
  const int n = d->n_out;
  int bn = 0;
  bool two = false;
  {
    const long long m_tiles = (long long)((d->W + d->box_w - 1) / d->box_w) * ((d->H + d->box_h - 1) / d->box_h) *
                              ((d->N + d->box_n - 1) / d->box_n);
    const long long k_iters = (long long)d->num_taps * (d->K / 64);
    const int cands[5] = {256, 160, 128, 64, 32};
    const int pair_mode = d->cta_pair == 1 ? 0 : (d->cta_pair == 2 ? 1 : g_pair_mode);
    const bool pair_ok = !d->b_batched && sms >= 2 && pair_mode != 0;
    if (d->cta_pair == 2 && !pair_ok) { set_last_error("tap_gemm: cta_pair=2 is not legal here (batched B or < 2 SMs)"); return G4_ERR_BAD_ARG; }
    if (d->tile_n != 0 && d->tile_n != 32 && d->tile_n != 64 && d->tile_n != 128 && d->tile_n != 160 && d->tile_n != 256) {
      set_last_error("tap_gemm: tile_n=%d (0, 32, 64, 128, 160 or 256)", d->tile_n); return G4_ERR_BAD_ARG;
    }
    double best = 0;
    for (int pass = 0; pass < 2; ++pass) {
      const bool tw = pass == 0;
      if (tw && !pair_ok) continue;
      if (!tw && pair_ok && pair_mode == 1) continue;
      for (int ci = 0; ci < 5; ++ci) {
        const int c = cands[ci];
        if (d->tile_n != 0 && c != d->tile_n) continue;
        if (d->act == G4_ACT_GEGLU && (c == 160 || c == 32 || n % c)) continue;
        if (c > 32 && n <= c / 2 && c != 64) continue;                 // do not pad tiny outputs to wide tiles
        const long long n_tiles = (n + c - 1) / c;
        const long long units = tw ? sms / 2 : sms;
        const long long tiles = (tw ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;
        const long long waves = (tiles + units - 1) / units;
        const double bytes = 16384.0 + (tw ? 64.0 : 128.0) * c;
        const double kstep = bytes / 45.0 > 2.0 * c ? bytes / 45.0 : 2.0 * c;
        const double main_clk = (double)k_iters * kstep;
        const double epi_clk = 300.0 + (d->act == G4_ACT_GEGLU ? 14.0 : 10.0) * c;
        const double cost = main_clk + (double)(waves - 1) * (main_clk > epi_clk ? main_clk : epi_clk) + epi_clk;
        if (bn == 0 || cost < best) { bn = c; best = cost; two = tw; }
      }
    }
    if (bn == 0) {
      if (d->tile_n != 0) { set_last_error("tap_gemm: tile_n=%d is not legal for n_out=%d act=%d", d->tile_n, n, d->act); return G4_ERR_BAD_ARG; }
      bn = (d->act == G4_ACT_GEGLU) ? 64 : 32;
    }
  }

  // Split-K: layers with few output tiles and a long reduction (the 5x8 level: 640 rows, K up to 23 040) leave most
  // SMs idle.  The K range is cut into `split` parts that run side by side; each CTA stores its raw fp32 accumulator
  // into a plane of the caller's workspace and splitk_reduce_kernel adds the planes IN ORDER and applies the epilogue
  // (deterministic; no atomics).  d->split_k: 0 = decide here, 1 = never, n = exactly n parts.
  int split = 1;
  {
    const long long m_tiles_s = (long long)((d->W + d->box_w - 1) / d->box_w) * ((d->H + d->box_h - 1) / d->box_h) *
                                ((d->N + d->box_n - 1) / d->box_n);
    const long long out_tiles = m_tiles_s * ((n + bn - 1) / bn);
    const long long k_it = (long long)d->num_taps * (d->K / 64);
    const long long rows_total = (long long)d->W * d->H * d->N;
    const bool can = !two && !d->b_batched && d->act != G4_ACT_GEGLU && d->workspace != nullptr;
    if (d->split_k < 0 || d->split_k > 64) { set_last_error("tap_gemm: split_k=%d", d->split_k); return G4_ERR_BAD_ARG; }
    if (d->split_k > 1) {
      if (!can) { set_last_error("tap_gemm: split_k needs a workspace, a single-CTA tile, no GEGLU and no batched B"); return G4_ERR_BAD_ARG; }
      split = d->split_k;
    } else if (d->split_k == 0 && can && out_tiles * 2 <= sms && k_it >= 16) {
      split = (int)(sms / out_tiles);
      if (split > k_it / 8) split = (int)(k_it / 8);       // at least 8 k-steps per part
      if (split > 16) split = 16;
    }
    if (split > 1) {
      const long long k_per = (k_it + split - 1) / split;
      split = (int)((k_it + k_per - 1) / k_per);            // no empty part
      if ((unsigned long long)split * rows_total * n * sizeof(float) > d->workspace_bytes) {
        if (d->split_k > 1) { set_last_error("tap_gemm: split_k=%d needs %lld workspace bytes", split, (long long)split * rows_total * n * 4); return G4_ERR_WORKSPACE; }
        split = 1;
      }
    }
    if (split < 1) split = 1;
  }

  CUtensorMap tmA, tmB, tmC;
  memset(&tmC, 0, sizeof(tmC));
  // bf16 outputs with a 16-byte aligned base and row pitch leave through TMA stores (tile-shaped box, 32 columns)
  const int n_store = d->act == G4_ACT_GEGLU ? d->n_out / 2 : d->n_out;
  const bool tma_store = !d->out_fp32 && (d->ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0) &&
                         n_store >= 32 && !g_no_tma_store && split == 1;
  if (tma_store) {
    uint64_t dims[4] = {(uint64_t)n_store, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t strides[3] = {(uint64_t)d->ldc * 2, (uint64_t)d->ldc * 2 * (uint64_t)d->W,
                           (uint64_t)d->ldc * 2 * (uint64_t)d->W * (uint64_t)d->H};
    uint32_t box[4] = {32, (uint32_t)d->box_w, (uint32_t)d->box_h, (uint32_t)d->box_n};
    int rc = make_tmap_bf16(&tmC, d->out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
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
    uint32_t box[3] = {64, (uint32_t)(two ? bn / 2 : bn), 1};
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
  a.tma_store = tma_store ? 1 : 0;
  a.split_k = split;
  a.split_stride = 0;
  ReduceArgs ra;
  if (split > 1) {   // partial mode: raw fp32 accumulators into the workspace planes, epilogue in the reduce pass
    const long long rows_total = (long long)d->W * d->H * d->N;
    ra.ws = reinterpret_cast<const float*>(d->workspace); ra.plane = rows_total * n; ra.split = split;
    ra.M = rows_total; ra.n_out = n; ra.out = d->out; ra.ldc = d->ldc; ra.out_fp32 = d->out_fp32;
    ra.alpha = d->alpha; ra.bias = d->bias; ra.row_bias = d->row_bias; ra.row_bias_ld = d->row_bias_ld;
    ra.rows_per_bias = a.rows_per_bias; ra.act = d->act; ra.residual = d->residual; ra.ldr = d->ldr;
    a.out = d->workspace; a.ldc = n; a.out_fp32 = 1; a.alpha = 1.0f; a.bias = nullptr; a.row_bias = nullptr;
    a.act = G4_ACT_NONE; a.residual = nullptr; a.split_stride = ra.plane;
  }

  int rc;
  if (two) {
    switch (bn) {
      case 32: rc = launch_tap_gemm<32, true>(tmA, tmB, tmC, a, sms, stream); break;
      case 64: rc = launch_tap_gemm<64, true>(tmA, tmB, tmC, a, sms, stream); break;
      case 128: rc = launch_tap_gemm<128, true>(tmA, tmB, tmC, a, sms, stream); break;
      case 160: rc = launch_tap_gemm<160, true>(tmA, tmB, tmC, a, sms, stream); break;
      default: rc = launch_tap_gemm<256, true>(tmA, tmB, tmC, a, sms, stream); break;
    }
  } else {
    switch (bn) {
      case 32: rc = launch_tap_gemm<32, false>(tmA, tmB, tmC, a, sms, stream); break;
      case 64: rc = launch_tap_gemm<64, false>(tmA, tmB, tmC, a, sms, stream); break;
      case 128: rc = launch_tap_gemm<128, false>(tmA, tmB, tmC, a, sms, stream); break;
      case 160: rc = launch_tap_gemm<160, false>(tmA, tmB, tmC, a, sms, stream); break;
      default: rc = launch_tap_gemm<256, false>(tmA, tmB, tmC, a, sms, stream); break;
    }
  }
  if (rc != G4_OK || split == 1) return rc;
  const long long work = ra.M * ((ra.n_out + 7) / 8);
  long long blocks = (work + 255) / 256;
  if (blocks > 4ll * sms) blocks = 4ll * sms;
  launch_pdl(splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, ra);
  return check_launch("tap_gemm split-K reduce");
}
