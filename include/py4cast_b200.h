/*
 * py4cast_b200 — C ABI of the B200 (sm_100a) kernels behind py4cast's
 * autoregressive-training hot path.
 *
 * Drop-in boundary: every entry point takes plain device pointers, sizes and a
 * cudaStream_t (as void*); no torch types.  The reference is pure Python on
 * PyTorch, so each entry point replaces a *group of ATen calls* made by the
 * reference (file:line cited per function, relative to the reference tree) and
 * is bound from Python with ctypes (py4cast_b200/_lib.py).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated otherwise;
 *   - `dtype`: P4C_F32 (float) or P4C_BF16 (__nv_bfloat16) = storage type of
 *     activation tensors; all arithmetic is fp32 (accumulation included);
 *   - return value: 0 on success, negative P4C_E* on error; the message of the
 *     last error of the calling thread is returned by p4c_last_error();
 *   - no entry point synchronises the stream or allocates memory; every one
 *     is CUDA-graph capturable; workspaces are caller-provided;
 *   - no atomics on floating-point data anywhere: results are bit-reproducible
 *     run to run for a fixed launch configuration.
 */
#ifndef PY4CAST_B200_H
#define PY4CAST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P4C_F32 0
#define P4C_BF16 1

#define P4C_OK 0
#define P4C_EINVAL (-1)   /* bad argument / unsupported shape            */
#define P4C_ECUDA (-2)    /* CUDA runtime error (launch failure, ...)     */
#define P4C_ENOSPACE (-3) /* caller-provided workspace too small          */

#define P4C_HIDDEN 64      /* hidden width of every fused MLP (hidden_dims) */
#define P4C_MAX_SRC 3      /* input segments of a fused MLP                 */
#define P4C_MAX_PREV 8     /* max input time steps of the AR concat         */

const char* p4c_last_error(void);
int p4c_version(void);
/* number of kernels launched by this library since load / last reset (host counter) */
int64_t p4c_launch_count(void);
void p4c_reset_launch_count(void);
/* sizeof() of the ABI structs as compiled (binding self-check): 0 p4c_src_t, 1 p4c_mlp_t,
 * 2 p4c_mlp_bwd_t, 3 p4c_ar_x_t, 4 p4c_adamw_state_t, 5 p4c_pack_t, 6 p4c_wgrad_t; -1 otherwise */
int p4c_sizeof(int which);

/* ------------------------------------------------------------------------
 * CSR construction (bit-exact w.r.t. a stable sort by key).
 * Replaces PyG's implicit COO handling inside mfai InteractionNet
 * (scatter over `edge_index[1]`).
 *   keys    [E]  int64, values in [0, num_keys)
 *   rowptr  [num_keys+1] int32 out
 *   perm    [E]  int32 out: edge ids grouped by key, ascending edge id inside a group
 *   scratch [E + num_keys + 1] int32 workspace
 * Set-up entry point: outside CUDA-graph capture it synchronises the stream once to validate the
 * keys and returns P4C_EINVAL when one lies outside [0, num_keys) (no kernel indexes with it).
 * ---------------------------------------------------------------------- */
int p4c_csr_build(const int64_t* keys, int64_t E, int32_t num_keys, int32_t* rowptr,
                  int32_t* perm, int32_t* scratch, void* stream);

/* int64 -> int32 narrowing copy of an index vector (edge_index rows). */
int p4c_index_to_i32(const int64_t* src, int32_t* dst, int64_t n, void* stream);

/* ------------------------------------------------------------------------
 * Deterministic segment sum:  out[b, s, :] (+)= scale_s * sum_{j in [rowptr[s], rowptr[s+1])} src[b, perm[j], :]
 * (perm == NULL  =>  rows of a segment are contiguous: src[b, j, :]).
 * Replaces scatter_add / index_add_ of InteractionNet.aggregate (K3) and the
 * transposed scatters of its backward.  Row width is P4C_HIDDEN.
 *   mean != 0  => scale_s = 1 / max(1, segment length)  (PyG "mean" aggregation)
 *   accumulate != 0 => out += result
 * ---------------------------------------------------------------------- */
int p4c_segment_sum(const void* src, int64_t src_batch_stride, const int32_t* rowptr,
                    const int32_t* perm, int32_t num_seg, int32_t batch, void* out,
                    int64_t out_batch_stride, int mean, int accumulate, int dtype,
                    void* stream);

/* ------------------------------------------------------------------------
 * Fused 2-layer MLP:  y = [LayerNorm]( W2 * SiLU(W1 * x + b1) + b2 ) [+ residual]
 * with x = concat of up to 3 (optionally gathered) row segments.
 * Replaces, per call, the ATen chain index_select x2 + cat + Linear + SiLU +
 * Linear + LayerNorm (+ add) of mfai make_mlp / InteractionNet (K1,K2,K4,K5,K6).
 * ---------------------------------------------------------------------- */
typedef struct {
  const void* ptr;      /* rows of `width` contiguous elements, row stride `ld` elements */
  const int32_t* index; /* optional gather index [rows] into this source (NULL = identity) */
  int64_t batch_stride; /* elements between samples (0 = shared by all samples)  */
  int32_t ld;
  int32_t width;
  /* Segment-sum source (forward only, 64-wide, tensor-core engine): when seg_rowptr != NULL, input
   * row r is  seg_scale[r] * sum_{j in [seg_rowptr[r], seg_rowptr[r+1])} ptr[b][seg_perm ? seg_perm[j] : j][:]
   * i.e. InteractionNet.aggregate (K3) fused into the prologue of the node MLP (K4); `index` must be
   * NULL.  seg_out (optional, [batch][full_rows][64], same dtype) receives the summed rows (they are
   * the saved activation the backward needs).                                                     */
  const int32_t* seg_rowptr;
  const int32_t* seg_perm;
  const float* seg_scale;
  void* seg_out;
} p4c_src_t;

typedef struct {
  int32_t rows;  /* rows per sample (E or N) */
  int32_t batch; /* samples                  */
  int32_t nsrc;  /* 1..3; if nsrc > 1 every width must be P4C_HIDDEN */
  int32_t k1;    /* sum of widths (input features of layer 1)        */
  int32_t out_features; /* <= 64; must be 64 when layer_norm          */
  int32_t layer_norm;
  int32_t dtype; /* storage type of sources, outputs, saved activations */
  int32_t full_rows; /* rows per sample of the output-like arrays (out, out2, residual, saved
                        activations, dy_b, dh1, dx): 0 = `rows`; > rows when this call covers a
                        row sub-range of bigger tensors (pointers pre-offset by the caller) */
  p4c_src_t src[P4C_MAX_SRC];
  const float* w1; /* [64][k1]  (nn.Linear layout: out, in) */
  const float* b1; /* [64]                                  */
  const float* w2; /* [out_features][64]                    */
  const float* b2; /* [out_features]                        */
  const float* gamma; /* [64] or NULL */
  const float* beta;  /* [64] or NULL */
  float ln_eps;
  /* optional P4C_PACK_BYTES image of the weights produced by p4c_pack_weights (bf16, laid out
   * for the tensor-core engine); NULL = the kernels convert w1/w2 themselves (slower set-up).
   * The image must have been packed from the current values of w1..beta.                       */
  const void* packed;
  /* tuning hint (tensor-core forward only): valid rows per 128-row CTA tile, one of 16/32/64/128
   * (0 = 128).  Smaller tiles spread long segment sums (hundreds of edges per row) over more CTAs. */
  int32_t tile_rows;
} p4c_mlp_t;

/* ------------------------------------------------------------------------
 * Weight images for the tensor-core engine.  One image per MLP:
 *   [W1 k-block 0..2 : 3 x 8 KB][W2 : 8 KB][W2^T : 8 KB][W1^T k-block 0..2 : 3 x 8 KB]
 *   [b1, b2, gamma, beta : 4 x 64 fp32]
 * every 8 KB block is a 64x64 bf16 matrix in the 128-byte-swizzled K-major shared-memory layout
 * tcgen05.mma reads, so a CTA stages it with plain 16-byte copies.  K-blocks follow the sources:
 * a source of width w contributes ceil(w/64) blocks (columns beyond w are zero).
 * p4c_pack_weights converts `n` MLPs in ONE launch from a DEVICE table of descriptors.
 * ---------------------------------------------------------------------- */
#define P4C_PACK_BYTES (9 * 8192 + 4 * 64 * 4)
typedef struct {
  const float* w1; const float* b1; const float* w2; const float* b2;
  const float* gamma; const float* beta; /* NULL -> 1 / 0 */
  void* dst;                               /* P4C_PACK_BYTES, 16-byte aligned */
  int32_t k1;
  int32_t out_features;
  int32_t nsrc;
  int32_t width[P4C_MAX_SRC];
} p4c_pack_t;
int p4c_pack_weights(const p4c_pack_t* table_dev, int32_t n, void* stream);

/*   residual  : optional [batch][rows][out_features] added to the output (K4 "rec + aggr_mlp(..)")
 *   out       : [batch][rows][out_features]
 *   out2      : optional, = src[0] row (identity index) + y   (K5 edge update "edge + msg")
 *   save_h1   : optional [batch][rows][64] pre-activation of layer 1 (for backward)
 *   save_stats: optional [batch][rows][2] float (mean, rstd of the LayerNorm input)
 */
/* Engine selection for the fused MLP: 1 (default) = tcgen05 tensor cores wherever the shape
 * allows (bf16 storage, 64-wide sources), 0 = CUDA-core fp32 FFMA everywhere.  The environment
 * variable P4C_ENGINE=simt sets the initial value to 0. */
void p4c_set_engine(int tensor_cores);
/* The persistent tensor-core kernels launch at most 2*148 - n CTAs (default n = 0: two per SM).
 * With n > 0 some SMs keep a free CTA slot, so that small kernels of ANOTHER stream (the mesh
 * levels of a second sample) run beside a large kernel instead of queueing behind it.  Changes the
 * number of weight-gradient partial blocks: set it before the first backward call. */
void p4c_set_cta_reserve(int n);
int p4c_get_cta_reserve(void);
/* The tensor-core forward has two kernels: a warp-specialised, software-pipelined one (one CTA per
 * SM: producer / MMA / epilogue warps, 3-stage gather ring) for calls of at least `n` 128-row tiles
 * and the monolithic two-CTAs-per-SM kernel below that (and for segment-sum sources, out_features
 * != 64).  Default n = 2 * 148 (environment variable P4C_WS_MIN_TILES); n <= 0 restores the default,
 * a huge n disables the pipelined kernel.  Same results up to fp32 summation order. */
void p4c_set_ws_min_tiles(int64_t n);
int64_t p4c_get_ws_min_tiles(void);
/* launches of the pipelined forward kernel so far (tests assert that it really ran) */
int64_t p4c_ws_launch_count(void);
/* developer aid: when set (16 x int64 device buffer), CTA 0 of the tensor-core forward kernel stores
 * the SM clock at the phases of its first tile (set-up, gather, GEMM 1, epilogue 1, ...). NULL = off. */
void p4c_set_trace_buffer(void* device_buffer);
int p4c_get_engine(void);

int p4c_mlp_forward(const p4c_mlp_t* m, const void* residual, void* out, void* out2,
                    void* save_h1, float* save_stats, void* stream);

/* p4c_mlp_forward with the aggregation of InteractionNet fused into the epilogue (K2+K3), for edge
 * sets whose edges are sorted by receiver with a uniform in-degree `group` (1, 2, 4 or 8, dividing
 * rows and full_rows; e.g. the 4 nearest mesh nodes of every grid node in m2g):
 *   aggr_out[b][r / group][:] = scale * sum of the `group` consecutive messages y[b][r..][:]
 * (fp32 sum of the un-rounded messages, fixed butterfly order).  `out` (the messages) may be NULL:
 * they are then never written to memory.  Tensor-core engine, bf16 storage, out_features = 64. */
int p4c_mlp_forward_aggregate(const p4c_mlp_t* m, void* out, void* out2, void* save_h1,
                              float* save_stats, void* aggr_out, int32_t group, float scale,
                              void* stream);

/* Backward.
 *   dy_a (+ dy_b): gradient w.r.t. y; dy_a may be gathered through dy_a_index
 *                  (rows -> row of dy_a, e.g. d_aggr[receiver[e]]) and scaled per
 *                  row by dy_a_scale[index] (mean aggregation); dy_b optional, identity rows.
 *   dh1_out      : optional [batch][rows][64] gradient w.r.t. layer-1 pre-activation
 *                  (consumed by p4c_segment_sum + p4c_rows_matmul for gathered sources)
 *   dx[i]        : optional gradient w.r.t. source i (identity-indexed sources only),
 *                  [batch][rows][width_i] with row stride dx_ld[i]; dx_accumulate[i] != 0 => +=
 *   wgrad_partials: workspace, p4c_mlp_wgrad_partial_floats(m) floats; reduced (in a fixed
 *                  order) into the g_* buffers ( += ) by the same call.
 */
typedef struct {
  const void* dy_a;
  const int32_t* dy_a_index;
  const float* dy_a_scale;
  int64_t dy_a_batch_stride;
  const void* dy_b;
  const void* h1;        /* saved by forward; may be NULL when p4c_mlp_bwd_caps(m) & 1 */
  const float* stats;    /* saved by forward (layer_norm only); NULL with h1 */
  void* dh1_out;
  void* dx[P4C_MAX_SRC];
  int32_t dx_ld[P4C_MAX_SRC];
  int32_t dx_accumulate[P4C_MAX_SRC];
  int32_t dx0_add_dy; /* 1: dx[0] += dy_b ; 2: dx[0] += dy_a + dy_b (src[0] is also the residual) */
  float* g_w1; float* g_b1; float* g_w2; float* g_b2; float* g_gamma; float* g_beta;
  float* wgrad_partials;
  int64_t wgrad_partials_floats;
  /* 0: partials are overwritten and reduced into g_* by this call (self-contained);
   * 1: partials are overwritten, the reduction is deferred to p4c_wgrad_reduce_batched;
   * 2: this call ADDS to the partials already in the workspace (same grid), reduction deferred.
   * Deferred modes need a workspace that belongs to this MLP alone (g_* are ignored).          */
  int32_t partials_mode;
  /* ---- outputs only the pipelined recomputing backward engine produces (see p4c_mlp_bwd_caps) ----
   * dx0_batch_sum: optional [rows][64]: sum over the samples of dh1[b] x W1[:, K-block of src[0]] — the
   *   gradient of a 64-wide identity-indexed src[0] that is SHARED by the batch (batch_stride 0: static
   *   edge embeddings), written (not accumulated).  Replaces a (B, E, 64) input gradient + batch sum, or a
 *   second pass over dh1: every CTA runs the samples of its row blocks back to back and accumulates the
 *   product in 64 TMEM columns (caps bit 2: src[0] as above and at least 148 row blocks of 128 rows).
   * dx_group_out: optional [batch][full_rows / dx_group][64], ACCUMULATED in place: the input gradient
   *   of gathered source `dx_group_src` when row r gathers source row r / dx_group (receiver-sorted
   *   edges of uniform in-degree dx_group = 1, 2, 4 or 8): out[b][r / group] += sum over the group's
   *   rows of dh1 x W1[:, block].  Replaces one p4c_seg_linear pass over dh1.                          */
  void* dx0_batch_sum;
  void* dx_group_out;
  int32_t dx_group;
  int32_t dx_group_src;
} p4c_mlp_bwd_t;

/* Capabilities of the backward engine that will serve the problem `m` (as the BACKWARD describes it:
 * segment-sum sources replaced by their saved sums), bit mask:
 *   1  recomputes h1 and the LayerNorm statistics from the sources: g->h1 / g->stats may be NULL, i.e.
 *      the forward need not save them (pass save_h1 = save_stats = NULL).  Not valid with dy_a_scale.
 *   2  produces dx0_batch_sum        4  produces dx_group_out
 * 0 = the classic engines (h1 and stats required, none of the extra outputs).  The pipelined engine
 * (one persistent CTA per SM: producer / MMA / two compute groups, 512 TMEM columns) takes bf16 calls
 * with LayerNorm, 64 outputs, packed weights and at least p4c_get_ws_bwd_min_tiles() 128-row tiles. */
int32_t p4c_mlp_bwd_caps(const p4c_mlp_t* m);
void p4c_set_ws_bwd_min_tiles(int64_t n); /* n <= 0: default (2 * 148; env P4C_WS_BWD_MIN_TILES) */
int64_t p4c_get_ws_bwd_min_tiles(void);
int64_t p4c_ws_bwd_launch_count(void);
int64_t p4c_ws_bwd_bsum_launch_count(void); /* launches that produced dx0_batch_sum */

/* number of per-CTA partial blocks p4c_mlp_backward writes for this problem (its grid) */
int32_t p4c_mlp_wgrad_blocks(const p4c_mlp_t* m);

/* Deferred weight-gradient reduction of `n` MLPs in ONE launch (DEVICE table): g_* += sum over the
 * per-CTA partial blocks, in a fixed order.  Any g_* pointer may be NULL.                        */
typedef struct {
  const float* partials;
  float* g_w1; float* g_b1; float* g_w2; float* g_b2; float* g_gamma; float* g_beta;
  int32_t nblocks;
  int32_t k1;
  int32_t out_features;
  int32_t _pad;
} p4c_wgrad_t;
int p4c_wgrad_reduce_batched(const p4c_wgrad_t* table_dev, int32_t n, void* stream);

/* out[b, r, :] (+)= ( scale[r] * sum_{j in seg r} src[b, perm ? perm[j] : j, :] ) x W[:, col0 : col0+64]
 * with W = w1 [64][k1]: segment sum + input-gradient GEMM of one 64-wide gathered source in one
 * kernel (replaces p4c_segment_sum + p4c_rows_matmul in the InteractionNet backward).  Tensor-core
 * engine, bf16 storage only (fp32 callers keep the two-call form).  `packed` (optional) is the weight
 * image of the MLP w1 belongs to and `kblock` the K-block of it that starts at column col0 (the
 * kernel converts w1 itself when packed is NULL).                                                */
int p4c_seg_linear(const void* src, int64_t src_batch_stride, const int32_t* rowptr,
                   const int32_t* perm, const float* scale, int32_t num_seg, int32_t batch,
                   const float* w1, int32_t k1, int32_t col0, const void* packed, int32_t kblock,
                   void* out, int64_t out_batch_stride, int accumulate, int32_t tile_rows /* as in
                   p4c_mlp_t */, int dtype, void* stream);

/* One or TWO segment sets over the same rows `src` in ONE launch (the InteractionNet backward needs the
 * gradient by receiver and by sender of the same dh1: reference autograd of `index_select` x 2,
 * SURVEY.md 3.2 K1 / K3 transposed).  part1 may be NULL (= p4c_seg_linear).  When both parts name the same
 * `out` (senders are receivers: same-level mesh InteractionNets) the two summed tiles are the two K-blocks
 * of one product and `out` is read-modify-written once (part1->accumulate must be 1; part0->accumulate says
 * whether the previous content of `out` is added); otherwise the parts are independent problems that share
 * the launch (CTAs split in proportion to their tiles).                                           */
typedef struct {
  const int32_t* rowptr;
  const int32_t* perm;
  const float* scale;
  int32_t num_seg;
  int32_t batch;
  int64_t src_batch_stride;
  int32_t col0;
  int32_t kblock;
  void* out;
  int64_t out_batch_stride;
  int32_t accumulate;
  int32_t tile_rows;
} p4c_seg_part_t;
int p4c_seg_linear_pair(const void* src, const float* w1, int32_t k1, const void* packed,
                        const p4c_seg_part_t* part0, const p4c_seg_part_t* part1, int dtype, void* stream);

int64_t p4c_mlp_wgrad_partial_floats(const p4c_mlp_t* m);
int p4c_mlp_backward(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g, void* stream);

/* out[b, r, :] (+)= in[b, r, :] (64 wide) x W[:, col0 : col0+64]  with W = w1 [64][k1]
 * i.e. the input-gradient of one 64-wide input segment from (segment-summed) dh1.
 * Used for the sender / receiver gradients of the edge MLP (linearity of dX = dh1 W1). */
int p4c_rows_matmul(const void* in, int64_t in_batch_stride, const float* w, int32_t k1,
                    int32_t col0, int32_t rows, int32_t batch, void* out,
                    int64_t out_batch_stride, int accumulate, int dtype, void* stream);

/* sum over the batch dimension: out[r, :] (+)= sum_b in[b, r, :]   (width elements per row) */
int p4c_batch_sum(const void* in, int64_t batch_stride, int32_t batch, int64_t n, void* out,
                  int accumulate, int dtype, void* stream);

/* ------------------------------------------------------------------------
 * AR step (reference py4cast/lightning.py:565-662).
 * ---------------------------------------------------------------------- */
typedef struct {
  int32_t batch;
  int32_t nodes;        /* N (grid points; H*W for conv layouts) */
  int32_t f_state;      /* F   */
  int32_t f_static;     /* F_s */
  int32_t f_forcing;    /* F_f */
  int32_t num_prev;     /* T_in */
  const float* prev[P4C_MAX_PREV];      /* prev[t]: [batch][nodes][F], oldest first  */
  int64_t prev_batch_stride[P4C_MAX_PREV];
  const float* stat;     /* [batch?][nodes][F_s] */
  int64_t stat_batch_stride; /* 0 if shared */
  const float* forcing;  /* [batch][nodes][F_f] for the step the x is built for */
  int64_t forcing_batch_stride;
  int32_t x_ld;          /* row stride of x (elements); 0 = dense (T_in*F + F_s + F_f).  Padding
                            columns are written as zeros (the tensor-core MLP reads them).       */
} p4c_ar_x_t;

/* x[b,n,:] = cat(prev[0..T_in-1], static, forcing)      (lightning.py:711-767 `_next_x`)
 * x: [batch][nodes][T_in*F + F_s + F_f] of `x_dtype`.  num_prev = 0 is allowed here (and only here): the
 * `downscaling_only` strategy feeds statics and forcings alone (lightning.py:759-765). */
int p4c_ar_concat(const p4c_ar_x_t* a, void* x, int x_dtype, void* stream);
/* 0: p4c_ar_concat / p4c_ar_step_forward / p4c_ar_step_backward use their scalar kernels only (the staged /
 * vectorised ones need nodes % 64 == 0 resp. nodes * F % 4 == 0 and 16-byte aligned arrays; same results bit for
 * bit: tests compare the two); default 1. */
void p4c_set_ar_vector_paths(int enabled);

/* new_state = border*truth + interior*(prev_last + y*std + mean)   (lightning.py:598-633)
 *   scale_y == 0      => predicted = prev_last + y                  (diff_ar, :623)
 *   truth/border NULL => new_state = predicted                      (inference or diff_ar)
 * new_state is written with batch stride new_state_batch_stride (a slot of the (B,T,N,F)
 * prediction slab, replacing torch.stack at :660).  If `next` != NULL the x of the NEXT
 * model call is produced in the same pass: next->prev[T_in-1] is ignored and replaced by
 * new_state.                                                                             */
int p4c_ar_step_forward(const void* y, int y_dtype, const float* prev_last,
                        int64_t prev_last_batch_stride, const float* truth,
                        int64_t truth_batch_stride, const float* border /*[nodes]*/,
                        const float* diff_std /*[F] or NULL*/, const float* diff_mean,
                        int32_t batch, int32_t nodes, int32_t f_state, float* new_state,
                        int64_t new_state_batch_stride, const p4c_ar_x_t* next, void* x_next,
                        int x_dtype, void* stream);

/* Backward of p4c_ar_step_forward, for the autograd node
 *     (y, prev[0..T_in-1]) -> (new_state, x_next).
 *   g_new   : gradient w.r.t. new_state from its other consumers ([batch][nodes][F], may be NULL)
 *   g_new2  : a second addend of the same gradient (NULL or [batch][nodes][F]): the input gradient the NEXT
 *             step's backward produced for this state, summed here instead of by a separate element-wise
 *             pass (torch autograd's accumulation of the two gradients of a state that feeds both the loss and
 *             the next step, reference lightning.py:598-662)
 *   g_xnext : gradient w.r.t. x_next ([batch][nodes][f_in] of x_dtype, may be NULL)
 *   G = (g_new + g_new2) + g_xnext[slot T_in-1];  dy = G*interior*(std);
 *   d_prev[T_in-1] = G*interior + g_xnext[slot T_in-2]   (prev_last is also copied into x_next)
 *   d_prev[t]      = g_xnext[slot t-1]   for 1 <= t <= T_in-2;   d_prev[0] unused (T_in >= 2).
 *   d_prev: HOST array of T_in device pointers (NULL entries are skipped); all outputs are
 *   written, not accumulated. */
int p4c_ar_step_backward(const float* g_new, int64_t g_new_batch_stride, const float* g_new2,
                         int64_t g_new2_batch_stride, const void* g_xnext, int x_dtype, const float* border, const float* diff_std,
                         int32_t batch, int32_t nodes, int32_t f_state, int32_t num_prev,
                         int32_t f_in /* row stride of g_xnext */, void* dy, int y_dtype,
                         float* const* d_prev, void* stream);

/* ------------------------------------------------------------------------
 * Losses (reference py4cast/losses.py:103-210).
 *   kind: 0 = MSELoss, 1 = L1Loss
 *   form: 0 = WeightedLoss        -> out [BT]      (losses.py:130-169)
 *         1 = ScaledLoss          -> out [BT][F]   (losses.py:186-210, sqrt iff MSE)
 *         2 = WeightedLoss, reduce_spatial_dim=False -> out [BT][N]  (losses.py:152-154)
 *   pred/target: [BT][N][F] fp32;  mask: same shape fp32 or NULL (= all ones, lightning.py:797)
 *   interior: [N] fp32 (any per-pixel weight);  w: [F]
 *   denom: num_interior - #{pixels masked in every sample/step/feature} — computed by
 *          p4c_loss_denominator into a device float (losses.py:156,164-167).
 *   partials: workspace of p4c_loss_partial_floats(...) floats.
 * ---------------------------------------------------------------------- */
int p4c_loss_denominator(const float* mask, int64_t bt, int32_t nodes, int32_t f,
                         float num_interior, float* denom_out, int32_t* scratch /*[nodes]*/,
                         void* stream);
int64_t p4c_loss_partial_floats(int64_t bt, int32_t nodes, int32_t f, int form);
int p4c_loss_forward(const float* pred, const float* target, const float* mask,
                     const float* interior, const float* w, const float* denom, int64_t bt,
                     int32_t nodes, int32_t f, int kind, int form, float* partials,
                     float* out, void* stream);
/* dpred = d(sum(gout * out))/dpred ; gout has the shape of `out`; `out` = forward result
 * (needed for form 1 + MSE).  dtarget (optional) receives -dpred-equivalent gradient. */
int p4c_loss_backward(const float* pred, const float* target, const float* mask,
                      const float* interior, const float* w, const float* denom,
                      const float* gout, const float* out, int64_t bt, int32_t nodes,
                      int32_t f, int kind, int form, float* dpred, float* dtarget,
                      void* stream);

/* -------------------------------------------------------------------------
 * External forcings of a sample window (input pipeline).
 * Replaces reference py4cast/forcingutils.py:61-90 (get_year_hour_forcing) + :93-132
 * (generate_toa_radiation_forcing) as assembled by datasets/base.py:233-274 (generate_forcings)
 * and broadcast / concatenated at base.py:509-515: writes channels ch0.. of the
 * (slots, nodes, ld) fp32 forcing tensor in one pass.
 *   lat, lon [nodes] degrees; hour [slots] = UTC hour of day of each time slot;
 *   sin_dec, cos_dec [slots] = sin / cos of the solar declination (radians) of each slot;
 *   datetime4 [slots][4] = (cos_hour, sin_hour, cos_doy, sin_doy) columns of get_year_hour_forcing,
 *   or NULL: then only the top-of-atmosphere radiation is written (1 channel instead of 5).
 * ---------------------------------------------------------------------- */
int p4c_forcing_fill(const float* lat, const float* lon, const float* hour, const float* sin_dec,
                     const float* cos_dec, const float* datetime4, int32_t slots, int32_t nodes,
                     float* out, int32_t ld, int32_t ch0, void* stream);

/* The whole forcing tensor of a batch in one pass (reference datasets/base.py:503-515: the NWP forcing
 * parameters followed by the five generated channels, concatenated per item and stacked by collate_fn,
 * base.py:173-195): out[slot][node][0..n_ship) = src[slot][node][:] converted to fp32 (src_dtype P4C_F32 or
 * P4C_BF16: the dtype the channels crossed PCIe in), out[..][n_ship..n_ship+4] = (cos_hour, sin_hour, cos_doy,
 * sin_doy, toa_radiation) as p4c_forcing_fill computes them (bit-identical), columns beyond n_ship+5 zero.
 * Rows are written whole (coalesced); ld >= n_ship + 5. */
int p4c_forcing_assemble(const void* src, int src_dtype, int32_t n_ship, const float* lat, const float* lon,
                         const float* hour, const float* sin_dec, const float* cos_dec, const float* datetime4,
                         int32_t slots, int32_t nodes, float* out, int32_t ld, void* stream);

/* -------------------------------------------------------------------------
 * Anomaly correlation coefficient (reference py4cast/metrics.py:407-420,
 * MetricACC.update): out[bt][f] = S0 / sqrt(S1 * S2) over the spatial points of one (sample, step),
 * S0 = sum (p - c_f)(t - c_f) m, S1 = sum ((p - c_f) m)^2, S2 = sum ((t - c_f) m)^2; mask optional
 * (NULL = ones).  pred / target / mask: fp32 (bt, nodes, f) contiguous; partials: workspace of
 * p4c_acc_partial_floats() floats.
 * ---------------------------------------------------------------------- */
int64_t p4c_acc_partial_floats(int64_t bt, int32_t nodes, int32_t f);
int p4c_acc_forward(const float* pred, const float* target, const float* mask,
                    const float* climate_means, int64_t bt, int32_t nodes, int32_t f,
                    float* partials, float* out, void* stream);

/* -------------------------------------------------------------------------
 * Power spectral density (reference py4cast/metrics.py:253-352 `dct_2d`, `dct_var`,
 * `radial_bin_dct`, `power_spectral_density`, called by MetricPSDK.add_psd :121-144 and
 * MetricPSDVar.update :186-213).  The reference's radial binning reads the flattened variance spectrum at
 * 2 r - 1, 2 r, 2 r + 1 (metrics.py:311-315): the first 2 Rmax + 1 entries of its first row and its last
 * entry.  This entry point produces exactly those, per (sample-step, feature) field, in one pass:
 *   x, mask (optional, NULL = ones): fp32 (bt, H, W, F) contiguous (features last);
 *   wy [H] = cos(pi (2 y + 1)(H - 1) / (2 H))  (last-row DCT weights, host-computed once);
 *   ctab [W][J + 1]: ctab[x][j] = cos(pi (2 x + 1) j / (2 W)) for j < J, ctab[x][J] = the same with
 *     j = W - 1 (host-computed once per grid shape, in fp64);
 *   spec (bt, F, J + 1): entries j < J = (1 / W^2) DCT2(x m)[0][j]^2, entry J = (1 / W^2) DCT2(x m)[H-1][W-1]^2;
 *   partials: workspace of p4c_psd_partial_floats() floats.
 * ---------------------------------------------------------------------- */
int64_t p4c_psd_partial_floats(int64_t bt, int32_t H, int32_t W, int32_t F);
int p4c_psd_first_row(const float* x, const float* mask, const float* wy, const float* ctab,
                      int64_t bt, int32_t H, int32_t W, int32_t F, int32_t J, float* partials,
                      float* spec, void* stream);
/* Radial profile from the spectrum entries above (reference metrics.py:311-321 radial_bin_dct with its flat
 * indexing, :343 batch mean): spec [B][F][2R + 1] -> prof [F][R] = mean_b spec[2r] + 0.5 mean_b spec[2r - 1 (or the
 * last entry for r = 0)] + 0.5 mean_b spec[2r + 1]; present [R] (bytes) marks the bins that hold at least one
 * pixel, the others are NaN (0 / 0 in the reference). */
int p4c_psd_profile(const float* spec, const unsigned char* present, int32_t B, int32_t F, int32_t R, float* prof,
                    void* stream);

/* -------------------------------------------------------------------------
 * Dataset statistics (reference py4cast/datasets/compute_dataset_stats.py:11-127).
 * Per (sample, feature) of x (B, X, F) fp32 — or of x - prev when prev != NULL (time differences of
 * compute_time_step_stats :88-127): out (B, F, 4) = NaN-ignoring mean over X (`nanmean`, :47 / :112),
 * NaN-ignoring mean of squares (:48 / :113), min with NaN -> +inf (:50), max with NaN -> -inf (:54).
 * fp64 accumulation in a fixed order.  partials: workspace of p4c_feature_stats_partial_bytes() BYTES.
 * ---------------------------------------------------------------------- */
int64_t p4c_feature_stats_partial_bytes(int32_t B, int64_t X, int32_t F);
int p4c_feature_stats(const float* x, const float* prev, int32_t B, int64_t X, int32_t F,
                      void* partials, float* out, void* stream);

/* Un-normalisation of predictions in place (reference py4cast/lightning.py:1162-1169): for every row
 * x[r][f] = x[r][f] * scale[f] + shift[f], rounded after the product and after the sum like the
 * reference's `*= std` followed by `+= mean`.  x: fp32 (rows, F) contiguous. */
int p4c_scale_shift(float* x, int64_t rows, int32_t F, const float* scale, const float* shift,
                    void* stream);

/* ------------------------------------------------------------------------
 * Optimiser: AdamW on one flat fp32 buffer (reference lightning.py:442-467:
 * torch.optim.AdamW + cosine-with-min-lr schedule with warm-up).
 * `state` is a device struct advanced by p4c_adamw_advance (one tiny kernel, so the
 * whole step is CUDA-graph replayable without host-side scalar updates).
 * ---------------------------------------------------------------------- */
typedef struct {
  int64_t step;        /* optimiser steps taken so far           */
  float lr;            /* learning rate for the current step     */
  float bias1;         /* 1 - beta1^step                         */
  float bias2;         /* 1 - beta2^step                         */
  float _pad;
} p4c_adamw_state_t;

int p4c_adamw_advance(p4c_adamw_state_t* state, float base_lr, float min_lr,
                      int64_t warmup_steps, int64_t total_steps, float beta1, float beta2,
                      void* stream);
int p4c_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   int64_t n, const p4c_adamw_state_t* state, float beta1, float beta2,
                   float eps, float weight_decay, float grad_scale, void* stream);

/* utility: y = x cast (fp32 <-> bf16), n elements */
int p4c_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* utility: dst[i] = a[i] + b[i] (fp32), used for gradient fan-in without ATen */
int p4c_add(const void* a, const void* b, void* dst, int64_t n, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PY4CAST_B200_H */
