/*
 * geo4d_b200 -- C ABI of the B200-native Geo4D inference hot path.
 *
 * The reference (jzr99/Geo4D @ 2e57abc) is pure Python/PyTorch and has NO
 * plugin / operator / FFI interface: its seams are Python
 * class paths named in configs/inference_geo4d.yaml and plain method calls.
 * Each entry point below therefore names the reference call site whose
 * arithmetic it replaces; the host-side mirrors of those Python seams live in
 * geo4d_b200/ (same class names, arguments and state-dict keys) and reach the
 * kernels through this library only.
 *
 * Conventions (all entry points):
 *   - pointers are DEVICE pointers owned by the caller (the PyTorch caching
 *     allocator on the Python side); the library never allocates or frees device
 *     memory and never synchronises;
 *   - work is enqueued on the caller-supplied stream (CUDA-graph capturable);
 *   - return value 0 on success, negative g4 error code otherwise, with a
 *     human readable message available from geo4d_last_error() (thread-local);
 *   - activations are bf16 "frames-major channels-last": row (n, y, x) of a
 *     feature map [N, H, W, C] is at ((n*H + y)*W + x) * ld + c;
 *   - fp32 is used for statistics, accumulators, latents and all geometry.
 */
#ifndef GEO4D_B200_H_
#define GEO4D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* g4_stream_t; /* cudaStream_t */

#define GEO4D_ABI_VERSION 3

int geo4d_abi_version(void);
const char* geo4d_last_error(void);
/* 1 if the current device is compute capability 10.x (tcgen05/TMEM present). */
int geo4d_device_supported(void);
/* Number of kernels this library has launched (or recorded into a CUDA graph) in this process. */
uint64_t geo4d_launch_count(void);
/* Debug aid: when non-null, every geo4d_tap_gemm launch writes up to 16 %globaltimer stamps per CTA into
 * buf[grid][16] (entry, setup done, first operands landed, first tile's MMAs issued, first accumulator
 * ready, first / last epilogue done, exit).  Pass NULL to switch it off (the default). */
void geo4d_debug_gemm_trace(void* buf);
/* Debug aid: non-zero forces the direct (per-thread) store epilogue instead of shared memory + TMA stores. */
void geo4d_debug_gemm_direct_store(int on);
/* Debug aid: -1 = cost model (default), 0 = never pair CTAs (cta_group::1 only), 1 = always pair when legal. */
void geo4d_debug_gemm_pair_mode(int mode);

/* ------------------------------------------------------------------------------------------------
 * Tap-GEMM on tcgen05 tensor cores (TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue).
 *
 *   out[(n,y,x), j] = epilogue( sum_{tap} sum_{c<K} A[n, y+dy[tap], x+dx[tap], c] * B[tap, j, c] )
 *
 * with zero padding outside [0,W)x[0,H).  One kernel covers every dense contraction of the path:
 *   - nn.Linear / 1x1 conv (to_q/to_k/to_v/to_out, proj_in/out, GEGLU, FF; attention.py:54-59,262-290,
 *     415-442; ae_modules.py nin_shortcut/AttnBlock q,k,v,proj_out):              num_taps = 1
 *   - 3x3 Conv2d, padding 1 (ResBlock in/out layers openaimodel3d.py:151-180, Upsample conv :98-106,
 *     VAE ResnetBlock/conv_in/conv_out ae_modules.py:189-248,604-659):             num_taps = 9
 *   - Conv3d (3,1,1), padding (1,0,0) (TemporalConvBlock openaimodel3d.py:255-279): num_taps = 3
 *     (view the tensor as W = h*w pixels, H = frames, N = batch)
 *   - batched matmul (VAE AttnBlock q k^T / p v, ae_modules.py:63-73):             b_batched = 1
 * K (= C) must be a multiple of 64 (pad the channel dim with zeros otherwise).
 * ---------------------------------------------------------------------------------------------- */
enum { G4_ACT_NONE = 0, G4_ACT_SILU = 1, G4_ACT_GEGLU = 2, G4_ACT_GELU = 3 /* exact-erf GELU (nn.GELU()) */ };

typedef struct {
  /* A operand, bf16 */
  const void* a;
  int K;                     /* channels (multiple of 64)                         */
  int W, H, N;               /* logical extents                                    */
  int64_t a_stride_w, a_stride_h, a_stride_n; /* in elements                       */
  int box_w, box_h, box_n;   /* M tile = box_w*box_h*box_n rows, <= 128            */
  int num_taps;              /* 1..9                                               */
  int tap_dx[9], tap_dy[9];
  /* B operand, bf16 [num_taps][n_out][K] (K contiguous); if b_batched, [N][n_out][K]
   * and the matrix used is selected by the row tile's n index (box_n must be 1). */
  const void* b;
  int n_out;
  int b_batched;
  /* output: bf16 (or fp32 if out_fp32) at out + row*ldc + col, row = (n*H + y)*W + x.
   * For G4_ACT_GEGLU the stored width is n_out/2 (B rows interleaved in blocks of 32:
   * 32 value rows followed by their 32 gate rows). */
  void* out;
  int64_t ldc;
  int out_fp32;
  float alpha;               /* accumulator scale applied first                    */
  const float* bias;         /* [n_out] fp32 or NULL                               */
  const float* row_bias;     /* fp32 [*, row_bias_ld]: added as row_bias[(row / rows_per_bias)][col] */
  int64_t row_bias_ld;
  int rows_per_bias;
  int act;
  const void* residual;      /* bf16, same row indexing with ldr, added last; may alias out */
  int64_t ldr;
  /* Tile overrides (ABI v2).  0 = let the library's cost model decide.  The result does not depend on them
   * (every output element sees the same k-step sequence); they exist so that a host can time the legal
   * configurations of a shape once and pin the fastest (geo4d_b200/ops.py: autotune). */
  int32_t tile_n;            /* 0 | 32 | 64 | 128 | 160 | 256 output columns per tile */
  int32_t cta_pair;          /* 0 auto | 1 single CTA (cta_group::1) | 2 CTA pair (cta_group::2, 256-row tiles) */
  /* Split-K (ABI v3): layers with few output tiles and a long reduction (5x8 latents, K up to 23 040) run their K
   * range in `split_k` parts side by side; partial fp32 tiles go to `workspace` and a second kernel adds them in
   * order and applies the epilogue (deterministic).  0 = library decides, 1 = never, n = n parts.  The workspace is
   * caller-owned scratch of at least split * rows * n_out * 4 bytes; NULL disables splitting.  The result is
   * independent of tile_n / cta_pair but (summation order) not of split_k. */
  int32_t split_k;
  int32_t reserved_;
  void* workspace;
  uint64_t workspace_bytes;
} g4_gemm_desc;

int geo4d_tap_gemm(const g4_gemm_desc* d, g4_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused softmax attention, head dim 64, non-causal, tcgen05 (S = QK^T and O = PV on tensor cores,
 * scores stay in TMEM/registers).  Replaces xformers.ops.memory_efficient_attention as called from
 * CrossAttention.efficient_forward (attention.py:146-209; einsum fallback :101-125).
 *   q   : bf16 [B*Lq rows, ldq]   head h at columns [h*64, h*64+64)
 *   k,v : bf16 [ceil(B/kv_batch_div)*Lk rows, ldkv]; batch entry b uses K/V block b / kv_batch_div
 *         (kv_batch_div = frames per clip for the text keys that repeat_interleave shares, 1 otherwise)
 *   out : bf16 [B*Lq rows, ldo];  accumulate != 0 adds to the existing contents (second
 *         cross-attention branch, attention.py:203-207).
 * ---------------------------------------------------------------------------------------------- */
int geo4d_attention(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                    int64_t ldo, int B, int H, int Lq, int Lk, int kv_batch_div, int accumulate, float scale,
                    g4_stream_t stream);

/* Both cross-attention branches of CrossAttention.efficient_forward in ONE launch (attention.py:166-207): text keys
 * (k, v: Lk = 77, shared by the frames of a clip through kv_batch_div) and image keys (k2, v2: Lk2 = 16 per frame)
 * are attended with the same Q tile and two independent softmaxes; out = softmax(QK^T)V + softmax(QK2^T)V2. */
int geo4d_cross_attention2(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int Lk,
                           int kv_batch_div, const void* k2, const void* v2, int64_t ldkv2, int Lk2,
                           int kv_batch_div2, void* out, int64_t ldo, int B, int H, int Lq, float scale,
                           g4_stream_t stream);

/* Temporal self-attention over <=16 frame tokens per (pixel, head) (CrossAttention.forward
 * attention.py:81-144 as used by TemporalTransformer :365-412).  Row of (b, t, p) = (b*T + t)*HW + p. */
int geo4d_temporal_attention(const void* q, const void* k, const void* v, int64_t ld, void* out, int64_t ldo,
                             int B, int T, int HW, int heads, float scale, g4_stream_t stream);

/* GroupNorm(32) [+SiLU] over `rows_per_stat` consecutive rows per statistic (basics.py:76-87,
 * openaimodel3d.py:151-155,175-180,256-266; attention.py:265,331; ae_modules.py:14-15).
 * `workspace` (geo4d_groupnorm_workspace_bytes) holds per-block partial sums and the completion tickets of
 * the statistics pass: its first 16 KiB must be ZERO before the first call (allocate it zero-filled once);
 * every call leaves them zero again, so one buffer serves any number of calls on a stream.  num_stats <= 4096. */
size_t geo4d_groupnorm_workspace_bytes(int num_stats, int rows_per_stat, int C);
/* Debug aid: non-zero keeps statistics and apply as two launches instead of one cooperative launch. */
void geo4d_debug_groupnorm_two_kernels(int on);
int geo4d_groupnorm_silu(const void* x, int64_t ldx, void* y, int64_t ldy, int num_stats, int rows_per_stat,
                         int C, const float* gamma, const float* beta, float eps, int apply_silu,
                         void* workspace, size_t workspace_bytes, g4_stream_t stream);
/* nn.LayerNorm over C per row (attention.py:229-231). */
int geo4d_layernorm(const void* x, int64_t ldx, void* y, int64_t ldy, int M, int C, const float* gamma,
                    const float* beta, float eps, g4_stream_t stream);

/* Layout / data movement (einops rearranges + torch.cat of openaimodel3d.py:588,627,633; ddpm3d.py:2541;
 * F.interpolate nearest x2 :98-106; stride-2 convs :66-77 and ae_modules.py:100-109 via im2col). */
int geo4d_bcthw_to_rows(const float* src0, int C0, const float* src1, int C1, int B, int T, int H, int W,
                        void* out_bf16, int Cpad, g4_stream_t stream);
int geo4d_rows_to_bcthw(const float* rows, int64_t ld, int C, int B, int T, int H, int W, float* out,
                        g4_stream_t stream);
int geo4d_concat_rows(const void* a, int64_t lda, int Ca, const void* b, int64_t ldb, int Cb, void* out,
                      int64_t rows, g4_stream_t stream);
int geo4d_upsample_nearest2x(const void* in, void* out, int N, int H, int W, int C, g4_stream_t stream);
int geo4d_im2col_3x3_s2(const void* in, void* out, int N, int H, int W, int C, int pad_before, int Ho, int Wo,
                        g4_stream_t stream);

/* DDIM update for the v-parameterisation (DDIMSampler.p_sample_ddim ddim.py:231-277;
 * predict_start/eps_from_z_and_v ddpm3d.py:278-290).  coef[step] = {sqrt(abar_t), sqrt(1-abar_t),
 * scale_prev/scale_t, sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma}; step = *step_idx (device). */
int geo4d_ddim_step(float* x, const float* v, float* pred_x0, const float* noise, const float* coef,
                    const int* step_idx, int64_t n, g4_stream_t stream);
int geo4d_advance_counter(int* counter, int delta, int modulo, g4_stream_t stream);
int geo4d_gather_row(const float* table, int64_t ld, const int* idx, float* out, int n, g4_stream_t stream);

/* VAE mid AttnBlock helpers (single head, d = C; ae_modules.py:53-78): row softmax of the fp32 score
 * matrix to bf16 probabilities, and a batched bf16 transpose out[b, c, r] = in[b, r, c] (to feed V^T as
 * the K-major B operand of the P V batched matmul). */
int geo4d_softmax_rows(const float* s, int64_t lds, void* p_bf16, int64_t ldp, int64_t rows, int cols,
                       g4_stream_t stream);
int geo4d_transpose_bf16(const void* in, int64_t ldin, void* out, int batch, int R, int C, g4_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Geometry: per-window post-processing and sliding-window alignment (all fp32, HBM-bound reductions).
 * ---------------------------------------------------------------------------------------------- */
/* infer_geo4d.py:447-487 fused: softplus(conf), sky / far masks, 1/conf, de-normalisation, (d+1)/2.
 * maps = [11][T][H][W] decoded channels (pts xyz, conf, ray dir xyz, ray moment xyz, inverse depth);
 * outputs pts [T*H*W][3], inv_conf / invdepth [T*H*W]; `valid` (optional) is AND-ed with ~invalid. */
int geo4d_postprocess_window(const float* maps, int64_t thw, float* pts, float* inv_conf, float* invdepth,
                             unsigned char* valid, float sky_value, float sky_eps, float far_value, float alpha,
                             float beta, int has_conf, g4_stream_t stream);
/* Pluecker ray map -> per-frame moments for the camera solve (utils/rays.py:301-367,387-433,579-595;
 * utils/normalize.py:25-51): out[t][18] = {sum(I - d d^T) (6 unique), sum (I - d d^T)(d x m) (3),
 * sum d_t (x) d_0 (9)} over the centre-cropped square; raydir/raymoment are [3][T][H][W]. */
int geo4d_raymap_moments(const float* raydir, const float* raymoment, int T, int H, int W, double* out,
                         g4_stream_t stream);
/* Weighted Umeyama moments = the reductions of roma.rigid_points_registration(x, y, weights=w1*w2,
 * compute_scaling=True) (init_im_poses.py:797-800).  pass 0: out = {sum w, sum w x(3), sum w y(3)};
 * pass 1 (means = {xm(3), ym(3)}): out = {sum w|x-xm|^2, sum w (y-ym)(x-xm)^T (9)}.  out: 10 doubles. */
/* Affine map of n_sets point sets [n_sets][pts_per_set][3] with one row-major 3x4 matrix each (mats [n_sets][12]):
 * mode 0 -> out [n_sets][pts_per_set][3] = A x + t (a window's registration applied to its point maps,
 * init_im_poses.py:317-330); mode 1 -> out [n_sets][pts_per_set] = third row only (camera-frame depth, :612-620). */
int geo4d_transform_points(const float* x, int n_sets, int64_t pts_per_set, const float* mats, float* out, int mode,
                           g4_stream_t stream);
int geo4d_umeyama_moments(const float* x, const float* y, const float* w1, const float* w2, int64_t n, int pass,
                          const double* means, double* out, g4_stream_t stream);
/* One fused iteration of the dense part of LightPointCloudGroupOptimizer.forward + backward + Adam on the
 * log-depth maps (optimizer_group.py:440-525; base_opt_group.py:593-626).  See csrc/align.cu for the
 * layout of scal / st and the reduced gradient outputs (all fp64, zeroed by the call). */
int geo4d_align_iter(float* logd, float* adam_m, float* adam_v, const float* pred, const float* weight,
                     const float* invd, const int* edge_ptr, const int* edge_idx, const float* poses,
                     const float* S, const float* scal, const float* invf, const int* it, const float* st,
                     double* gpose, double* gS, double* gscal, double* gst, int N, int G, int HW, int W,
                     int group_size, int max_edges_per_image, g4_stream_t stream);
/* One Adam iteration of the LAD scale/shift fit min sum|s x + t - y| for G windows at once
 * (absolute_value_scaling2 depth_eval.py:112-145).  state[g] = {s, t, m_s, v_s, m_t, v_t, prev_loss,
 * step, done}; acc = 4*G doubles of scratch (3 sums + an arrival ticket per window), zero before the first call. */
int geo4d_lad_step(const float* x, const float* y, int64_t n_per_group, int G, float* state, double* acc, float lr,
                   float tol, g4_stream_t stream);
/* HOST function (no CUDA): SQPnP from the 41 moments of one frame for focal f (the 9-unknown problem behind
 * cv2.solvePnPRansac(flags=SOLVEPNP_SQPNP) in fast_pnp, init_im_poses.py:824-865).  Writes the world-to-camera
 * rotation (row-major 3x3) and translation; returns 1 on success, 0 if there is no valid solution. */
int geo4d_sqpnp_from_moments(const double* mom, double f, double* R_out, double* t_out);
/* n problems at once on up to `threads` host threads; mom [n][41], R_out [n][9], t_out [n][3], ok_out [n]. */
int geo4d_sqpnp_from_moments_batch(const double* mom, const double* f, int n, double* R_out, double* t_out,
                                   int* ok_out, int threads);
/* The whole fit (up to `iters` iterations of geo4d_lad_step, same arithmetic and early exit) in one cooperative
 * launch with a per-window grid barrier; needs G <= number of SMs.  acc: geo4d_lad_fit_workspace_doubles(G) doubles of
 * scratch (per window: the published (s, t) words, an arrival ticket and one row of partial sums per block; the call
 * clears it).  The per-block sums are folded in block order: the fit is bit-reproducible. */
size_t geo4d_lad_fit_workspace_doubles(int G);
int geo4d_lad_fit(const float* x, const float* y, int64_t n_per_group, int G, float* state, double* acc, float lr,
                  float tol, int iters, g4_stream_t stream);
/* delta<1.25 accuracy of s*x+t vs y under (w>0.5 & x>0.05 & y>0) (depth_eval.py:296-317): out[g] = {ok, n}. */
int geo4d_delta125(const float* x, const float* y, const float* w, int64_t n_per_group, int G, const float* st,
                   int st_stride, double* out, g4_stream_t stream);

/* Initialisation solvers of the alignment on the GPU (the reference runs them on the CPU through
 * cv2.solvePnPRansac(SOLVEPNP_SQPNP) and scipy least_squares; init_im_poses.py:824-865,
 * utils/geometry.py:162-270).  The kernels reduce the (focal-independent) SQPnP moments and the shift/focal
 * sums; the 9x9 / 1-D solves stay on the host (geo4d_b200/init_solvers.py).  See csrc/align.cu for layouts. */
int geo4d_pnp_moments(const float* pts, const float* conf, int F, int HW, int W, float cx, float cy,
                      const float* gate, int C, float thr_px, double* out, g4_stream_t stream);
int geo4d_shift_focal_sums(const float* pts, const float* conf, int G, int HW, int W, int H, const float* shift,
                           float zoff, double* out, g4_stream_t stream);

/* O(N + G) part of one alignment step: chain rule to the reference's pose / scale / focal parametrisations
 * (base_opt_group.py:260-320, optimizer_group.py:193-198), temporal-smoothing and trajectory-prior terms
 * (optimizer_group.py:492-542), torch.optim.Adam on every small parameter, and the refreshed matrices for the
 * next geo4d_align_iter.  adam: geo4d_align_small_adam_floats(N, G) floats, zero-initialised. */
size_t geo4d_align_small_adam_floats(int N, int G);
int geo4d_align_small_step(float* im_poses, float* im_focal, float* pw_poses, float* s_depth, float* t_depth,
                           float* ta_poses, float* adam, const double* gpose, const double* gS, const double* gscal,
                           const double* gst, const float* traj, const int* e_img, const int* edge_ptr,
                           const int* edge_idx, const float* valid_traj, const float* scal, const int* it,
                           float* poses_out, float* S_out, float* invf_out, float* st_out, int N, int G,
                           int group_size, int start_b, float temporal_smoothing_weight, float translation_weight,
                           float base_scale, float focal_break, g4_stream_t stream);

/* The whole optimisation loop of LightPointCloudGroupOptimizer.compute_global_alignment for iterations
 * [it0, it1) (base_opt_group.py:553-626 driving optimizer_group.py:440-525) in ONE persistent cooperative
 * launch: dense per-pixel part (same arithmetic as geo4d_align_iter) -> grid barrier -> deterministic fold of
 * the per-unit partial sums -> (multi-GPU: images [n_lo, n_hi) of this rank; every rank's record is stored
 * straight into each peer's receive buffer over NVLink and summed in rank order) -> the O(N + G) step of
 * geo4d_align_small_step -> grid barrier.  Replaces (it1 - it0) x {geo4d_align_iter, geo4d_align_small_step}.
 * All pointers are device pointers owned by the caller; peer_rec / peer_flag are peer-mapped addresses
 * (cudaIpc / symmetric memory) of every rank's receive buffer ([2][world][rec_doubles] doubles) and flag
 * array ([world] uint64), world == 1 ignores them.  Flags carry flag_base + it + 1 and must grow
 * monotonically over the calls that share the buffers.  part: geo4d_align_loop_part_floats(n_hi - n_lo, chunks)
 * floats; bar: two zero-initialised uint32 (kept between calls).  HW must be a multiple of 4. */
typedef struct g4_align_loop_desc {
  float* logd; float* adam_m; float* adam_v;        /* [N][HW] log-depth and its Adam moments */
  const float* pred; const float* weight; const float* invd;   /* [E][HW][3], [E][HW], [E][HW] | NULL */
  const int* edge_ptr; const int* edge_idx;         /* image -> incident (window, frame) edges, CSR */
  const float* scal;                                /* [iters][8] per-iteration constants (see csrc/align.cu) */
  float* poses; float* S; float* invf; float* st;   /* [N][12], [G][12], [1], [G][3]: current matrices (in/out) */
  double* gpose; double* gS; double* gscal; double* gst;   /* [N][12], [G][12], [3], [G][2]: totals of the last iteration */
  float* part; unsigned int* bar;
  float* im_poses; float* im_focal; float* pw_poses; float* s_depth; float* t_depth; float* ta_poses;   /* parameters */
  float* adam_small;                                /* geo4d_align_small_adam_floats(N, G) floats */
  const float* traj; const int* e_img; const float* valid_traj;   /* [E][16], [E], [G] */
  int N, G, HW, W, group_size, max_edges_per_image;
  int n_lo, n_hi, chunks, it0, it1, start_b;
  float temporal_smoothing_weight, translation_weight, base_scale, focal_break;
  int world, rank;
  int img_lo[17];                                   /* image partition: rank r owns [img_lo[r], img_lo[r+1]) */
  int rec_doubles;                                  /* geo4d_align_loop_record_doubles(max images per rank, G) */
  void* peer_rec[16]; void* peer_flag[16];
  unsigned long long flag_base;
  void* debug_ns;                                   /* optional: 8 uint64, CTA 0's accumulated ns per stage (NULL = off) */
} g4_align_loop_desc;
size_t geo4d_align_loop_part_floats(int n_images_local, int chunks);
int geo4d_align_loop_record_doubles(int max_images_per_rank, int G);
int geo4d_align_loop_chunks(int n_images_local, int HW);
int geo4d_align_loop(const g4_align_loop_desc* d, g4_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GEO4D_B200_H_ */
