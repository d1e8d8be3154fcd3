// Fused softmax attention on tcgen05 tensor cores (head dim 64, non-causal), the "one true dense
// contraction" of the U-Net: replaces xformers.ops.memory_efficient_attention as called by
// CrossAttention.efficient_forward (attention.py:146-209) for spatial self-attention and for the two
// cross-attention branches (text keys, per-frame image keys; the second branch is summed into the first
// with `accumulate`, attention.py:203-207).  QK-scale, softmax and PV are fused; scores never leave the SM.
//
//  grid : one CTA per (batch, head, 128-query tile); two CTAs co-reside per SM (112 KB smem, 256 TMEM
//         columns each) so one CTA's softmax overlaps the other's MMAs.
//  warp 0 lane 0 : TMA producer  (Q once; K_j, V_j double buffered; 4-D maps read heads in place from the
//                  token-major [rows, heads*64] projections -- no head-split copies)
//  warp 1 lane 0 : tcgen05.mma   S = Q K_j^T  (M128 N128 K64, both K-major)        -> TMEM cols [0,128)
//                                O += P_j V_j (M128 N64 K128, P from smem K-major,
//                                              V MN-major straight from its TMA tile) -> TMEM cols [128,192)
//  warps 2..5    : one query row per thread: tcgen05.ld S, running max / sum in fp32, exp2 with the
//                  1/sqrt(d)*log2(e) scale folded in, rescale O in TMEM only when the max moved,
//                  write P (bf16, 128B-swizzled) to smem, final O / l -> bf16 store.
#include "common.cuh"
#include "geo4d_b200.h"

namespace g4 {

struct AttnArgs {
  int B, H, Lq, Lk;
  int n_qtiles, n_kvtiles;
  int kv_batch_div;  // K/V batch index = b / kv_batch_div (text keys shared by the frames of a clip)
  // second, independently normalised key/value set (image cross-attention, attention.py:203-207): its tiles follow the
  // first set's in the same CTA, O2 accumulates in its own TMEM columns and the epilogue stores O1/l1 + O2/l2
  int Lk2, n_kvtiles2, kv_batch_div2;
  int accumulate;
  float scale_log2;  // scale * log2(e)
  void* out;
  long long ldo;
};

constexpr int AT_Q_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int AT_KV_BYTES = 128 * 64 * 2;  // 16 KB per K or V tile
constexpr int AT_P_BYTES = 128 * 128 * 2;  // 32 KB
constexpr int AT_SMEM = AT_Q_BYTES + 4 * AT_KV_BYTES + AT_P_BYTES;  // 112 KB

template <bool TWO_SETS>
__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK2,
                const __grid_constant__ CUtensorMap tmV2, const AttnArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[13];
  __shared__ uint32_t tmem_slot;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_Q_BYTES;                    // [2]
  uint8_t* sV = smem + AT_Q_BYTES + 2 * AT_KV_BYTES;  // [2]
  uint8_t* sP = smem + AT_Q_BYTES + 4 * AT_KV_BYTES;
  uint64_t* q_full = &bars[0];
  uint64_t* k_full = &bars[1];   // [2]
  uint64_t* k_empty = &bars[3];  // [2]
  uint64_t* v_full = &bars[5];   // [2]
  uint64_t* v_empty = &bars[7];  // [2]
  uint64_t* s_full = &bars[9];
  uint64_t* p_ready = &bars[10];
  uint64_t* pv_done = &bars[11];
  uint64_t* s_free = &bars[12];  // the softmax warps hold S_j in registers: QK_{j+1} may overwrite it

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % args.n_qtiles;
  const int bh = blockIdx.x / args.n_qtiles;
  const int h = bh % args.H;
  const int b = bh / args.H;
  const int bkv = b / args.kv_batch_div;
  const int nkv1 = args.n_kvtiles;
  const int nkv = TWO_SETS ? nkv1 + args.n_kvtiles2 : nkv1;   // tiles j >= nkv1 belong to the second key/value set

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    if (TWO_SETS) { tma_prefetch_desc(&tmK2); tma_prefetch_desc(&tmV2); }
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(pv_done, 1);
    mbar_init(s_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_grid_sync();
  const uint32_t tS = tmem_base;        // 128 columns
  const uint32_t tO = tmem_base + 128;  // 64 columns (first set), + 64 columns at 192 for the second set

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, AT_Q_BYTES);
      tma_load_4d(sQ, &tmQ, q_full, 0, h, qt * 128, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const bool second = TWO_SETS && j >= nkv1;
        const int row0 = (second ? j - nkv1 : j) * 128;
        const int bb = second ? b / args.kv_batch_div2 : bkv;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], AT_KV_BYTES);
        tma_load_4d(sK + st * AT_KV_BYTES, second ? &tmK2 : &tmK, &k_full[st], 0, h, row0, bb);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], AT_KV_BYTES);
        tma_load_4d(sV + st * AT_KV_BYTES, second ? &tmV2 : &tmV, &v_full[st], 0, h, row0, bb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        const uint32_t aK = smem_u32(sK + st * AT_KV_BYTES);
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tS, make_sw128_desc(aQ + kk * 32, 16, 1024), make_sw128_desc(aK + kk * 32, 16, 1024), idesc_qk,
                  kk != 0);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
      };
      issue_qk(0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const uint32_t aV = smem_u32(sV + st * AT_KV_BYTES);
        // S_{j+1} = Q K_{j+1}^T as soon as the softmax warps have S_j in registers: it runs under softmax(j)
        if (j + 1 < nkv) {
          mbar_wait(s_free, j & 1);
          issue_qk(j + 1);
        }
        // O += P_j V_j
        mbar_wait(&v_full[st], ph);
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const uint32_t tOj = (TWO_SETS && j >= nkv1) ? tO + 64 : tO;          // each key/value set has its own accumulator
        const int jj = (TWO_SETS && j >= nkv1) ? j - nkv1 : j;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tOj, make_sw128_desc(aP + kb * 16384 + kk * 32, 16, 1024),
                    make_sw128_desc(aV + kb * 8192 + kk * 2048, 1024, 1024), idesc_pv, (jj | kb | kk) != 0);
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
    }
    __syncwarp();
  } else {
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const float sl2 = args.scale_log2;
    float m = -INFINITY, l = 0.f, l_first = 1.f;
    for (int j = 0; j < nkv; ++j) {
      if (TWO_SETS && j == nkv1 && j > 0) { l_first = l; m = -INFINITY; l = 0.f; }   // second set: a softmax of its own
      const bool second = TWO_SETS && j >= nkv1;
      const int jj = second ? j - nkv1 : j;
      const uint32_t tOj = second ? tO + 64 : tO;
      const int kv_valid = min(128, (second ? args.Lk2 : args.Lk) - jj * 128);  // keys of this tile that exist
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // the whole score row goes to registers with ONE TMEM round trip; S is then free for QK_{j+1}
      uint32_t sv[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tS + lane_off + c * 32, sv[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);
      float mx = -INFINITY;
      if (kv_valid == 128) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sv[c][i]));
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(sv[c][i]));
      }
      const float m_new = fmaxf(m, mx);
      const float alpha = (m == -INFINITY) ? 0.f : ex2_ftz((m - m_new) * sl2);
      const float mb = m_new * sl2;
      // the previous PV must have retired before O is rescaled / P is overwritten
      mbar_wait(pv_done, (j & 1) ^ 1);
      tc_fence_after();
      if (jj > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tOj + lane_off + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(tOj + lane_off + c * 32, v);
        }
        tmem_st_wait();
      }
      // p = exp2(s*sl2 - m*sl2), row sum, bf16 P tile (K-major, 128B swizzle) to smem
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint8_t* prow = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          float p[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = c * 32 + qq * 8 + i;
            const float e = ex2_ftz(fmaf(__uint_as_float(sv[c][qq * 8 + i]), sl2, -mb));
            p[i] = (kv_valid == 128 || col < kv_valid) ? e : 0.f;
            rs += p[i];   // fp32 row sum over the unrounded probabilities (as the flash kernels the reference calls do)
          }
          uint4 w;
          w.x = pack_bf16x2(p[0], p[1]);
          w.y = pack_bf16x2(p[2], p[3]);
          w.z = pack_bf16x2(p[4], p[5]);
          w.w = pack_bf16x2(p[6], p[7]);
          const int chunk = ((c & 1) * 4 + qq) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = w;
        }
      }
      l = l * alpha + rs;
      m = m_new;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_ready);
    }
    // epilogue: O / l -> bf16
    mbar_wait(pv_done, (nkv - 1) & 1);
    tc_fence_after();
    const int lq = qt * 128 + r;
    constexpr bool two = TWO_SETS;
    const float inv_l = two ? 1.0f / l_first : 1.0f / l;
    const float inv_l2 = 1.0f / l;
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(args.out) + ((long long)b * args.Lq + lq) * args.ldo + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32], v2[32];
      tmem_ld32(tO + lane_off + c * 32, v);
      if (two) tmem_ld32(tO + 64 + lane_off + c * 32, v2);
      tmem_ld_wait();
      if (lq < args.Lq) {
        uint4* o4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            o[i] = __uint_as_float(v[8 * qq + i]) * inv_l;
            if (two) o[i] = fmaf(__uint_as_float(v2[8 * qq + i]), inv_l2, o[i]);
          }
          if (args.accumulate) {
            const uint4 pr = o4[qq];
            float2 f;
            f = unpack_bf16x2(pr.x); o[0] += f.x; o[1] += f.y;
            f = unpack_bf16x2(pr.y); o[2] += f.x; o[3] += f.y;
            f = unpack_bf16x2(pr.z); o[4] += f.x; o[5] += f.y;
            f = unpack_bf16x2(pr.w); o[6] += f.x; o[7] += f.y;
          }
          uint4 w;
          w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
          w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
          o4[qq] = w;
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace g4

using namespace g4;

static int attention_impl(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int Lk, int kv_batch_div,
                          const void* k2, const void* v2, int64_t ldkv2, int Lk2, int kv_batch_div2, void* out, int64_t ldo,
                          int B, int H, int Lq, int accumulate, float scale, cudaStream_t stream) {
  if (!q || !k || !v || !out) { set_last_error("attention: null pointer"); return G4_ERR_BAD_ARG; }
  if (B < 1 || H < 1 || Lq < 1 || Lk < 1 || kv_batch_div < 1) { set_last_error("attention: bad sizes B=%d H=%d Lq=%d Lk=%d", B, H, Lq, Lk); return G4_ERR_BAD_ARG; }
  if (ldq % 8 || ldkv % 8 || ldo % 8 || ((uintptr_t)q & 15) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15) ||
      ((uintptr_t)out & 15)) {
    set_last_error("attention: pointers must be 16-byte aligned and leading dims multiples of 8"); return G4_ERR_BAD_ARG;
  }
  const bool two = k2 != nullptr;
  if (two && (!v2 || Lk2 < 1 || kv_batch_div2 < 1 || ldkv2 % 8 || ((uintptr_t)k2 & 15) || ((uintptr_t)v2 & 15))) {
    set_last_error("attention: bad second key/value set"); return G4_ERR_BAD_ARG;
  }
  CUtensorMap tmQ, tmK, tmV, tmK2, tmV2;
  {
    uint64_t dims[4] = {64, (uint64_t)H, (uint64_t)Lq, (uint64_t)B};
    uint64_t str[3] = {128, (uint64_t)ldq * 2, (uint64_t)ldq * 2 * (uint64_t)Lq};
    uint32_t box[4] = {64, 1, 128, 1};
    int rc = make_tmap_bf16(&tmQ, q, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {64, (uint64_t)H, (uint64_t)Lk, (uint64_t)((B + kv_batch_div - 1) / kv_batch_div)};
    uint64_t str[3] = {128, (uint64_t)ldkv * 2, (uint64_t)ldkv * 2 * (uint64_t)Lk};
    uint32_t box[4] = {64, 1, 128, 1};
    int rc = make_tmap_bf16(&tmK, k, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmV, v, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    tmK2 = tmK; tmV2 = tmV;
  }
  if (two) {
    uint64_t dims[4] = {64, (uint64_t)H, (uint64_t)Lk2, (uint64_t)((B + kv_batch_div2 - 1) / kv_batch_div2)};
    uint64_t str[3] = {128, (uint64_t)ldkv2 * 2, (uint64_t)ldkv2 * 2 * (uint64_t)Lk2};
    uint32_t box[4] = {64, 1, 128, 1};
    int rc = make_tmap_bf16(&tmK2, k2, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmV2, v2, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  AttnArgs a;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk;
  a.n_qtiles = (Lq + 127) / 128;
  a.n_kvtiles = (Lk + 127) / 128;
  a.kv_batch_div = kv_batch_div; a.accumulate = accumulate;
  a.Lk2 = two ? Lk2 : 0; a.n_kvtiles2 = two ? (Lk2 + 127) / 128 : 0; a.kv_batch_div2 = two ? kv_batch_div2 : 1;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.out = out; a.ldo = ldo;
  cudaError_t e = two ? cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)
                      : cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);   // per device, cheap
  if (e != cudaSuccess) { set_last_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return G4_ERR_CUDA; }
  const long long grid = (long long)a.n_qtiles * B * H;
  if (grid > 2147483647ll) { set_last_error("attention: grid too large"); return G4_ERR_UNSUPPORTED; }
  if (two) launch_pdl(attn_fwd_kernel<true>, dim3((int)grid), dim3(192), AT_SMEM, stream, tmQ, tmK, tmV, tmK2, tmV2, a);
  else launch_pdl(attn_fwd_kernel<false>, dim3((int)grid), dim3(192), AT_SMEM, stream, tmQ, tmK, tmV, tmK2, tmV2, a);
  return check_launch("attention");
}

extern "C" int geo4d_attention(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                               int64_t ldo, int B, int H, int Lq, int Lk, int kv_batch_div, int accumulate, float scale,
                               g4_stream_t stream_) {
  return attention_impl(q, ldq, k, v, ldkv, Lk, kv_batch_div, nullptr, nullptr, 0, 0, 1, out, ldo, B, H, Lq, accumulate, scale,
                        reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int geo4d_cross_attention2(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int Lk,
                                      int kv_batch_div, const void* k2, const void* v2, int64_t ldkv2, int Lk2,
                                      int kv_batch_div2, void* out, int64_t ldo, int B, int H, int Lq, float scale,
                                      g4_stream_t stream_) {
  if (!k2 || !v2) { set_last_error("cross_attention2: second key/value set is null"); return G4_ERR_BAD_ARG; }
  return attention_impl(q, ldq, k, v, ldkv, Lk, kv_batch_div, k2, v2, ldkv2, Lk2, kv_batch_div2, out, ldo, B, H, Lq, 0, scale,
                        reinterpret_cast<cudaStream_t>(stream_));
}
