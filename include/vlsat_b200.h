/*
 * vlsat_b200 - C ABI of the B200-native VL-SAT hot path (libvlsat_b200.so, sm_100a only).
 *
 * Every entry point takes raw DEVICE pointers + sizes + the cudaStream_t to launch on (passed as
 * void*), launches asynchronously, owns no persistent state, never synchronises, never throws.
 * Return value: 0 = ok, otherwise a vlsat_status code (vlsat_error_string() decodes it).
 * All floating-point tensors are fp32, row-major, innermost dimension contiguous unless a leading
 * dimension (ld*) argument says otherwise. Index tensors are int64 exactly as the reference passes
 * them (PyG requires int64 edge_index); derived CSR arrays are int32.
 *
 * Each declaration cites the reference code (paths relative to the reference root) it replaces.
 */
#ifndef VLSAT_B200_H
#define VLSAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    VLSAT_OK = 0,
    VLSAT_ERR_INVALID_ARG = 1,   /* null pointer / negative size / inconsistent dims               */
    VLSAT_ERR_UNSUPPORTED = 2,   /* shape outside what the kernels are built for (loud, no fallback) */
    VLSAT_ERR_LAUNCH = 3,        /* cudaGetLastError() after launch was not cudaSuccess             */
    VLSAT_ERR_WORKSPACE = 4      /* workspace_bytes too small                                       */
} vlsat_status;

#define VLSAT_ACT_NONE 0
#define VLSAT_ACT_RELU 1
#define VLSAT_ACT_SIGMOID 2

#define VLSAT_AGGR_MAX 0
#define VLSAT_AGGR_ADD 1
#define VLSAT_AGGR_MEAN 2

int vlsat_version(void);
const char* vlsat_error_string(int status);
/* Name of the GEMM engine the library was built with ("tcgen05-bf16x3", "simt-fp32", ...). */
const char* vlsat_gemm_engine(void);
/* Number of kernel launches issued by this process through the library so far (bench bookkeeping). */
int64_t vlsat_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * A1 / A3  PointNetfeat.forward  (src/model/model_utils/network_PointNet.py:121-176, batch_norm=False,
 *          input_transform=False, feature_transform=False, global_feat=True)
 *   out[o,c] = max_p ReLU(b3[c] + W3[c,:] . ReLU(b2 + W2 . ReLU(b1 + W1 . x[o,:,p])))
 *   x [n_obj, c_in, n_pts] channels-first; W1 [c1,c_in], W2 [c2,c1], W3 [c_out,c2] (Conv1d k=1 weights
 *   with the trailing 1 squeezed). argmax (nullable) [n_obj, c_out] int32 = arg max point (for backward).
 * ---------------------------------------------------------------------------------------------- */
int vlsat_pointnet_fwd(const float* x, int64_t n_obj, int c_in, int64_t n_pts,
                       const float* w1, const float* b1, int c1,
                       const float* w2, const float* b2, int c2,
                       const float* w3, const float* b3, int c_out,
                       float* out, int32_t* argmax, void* stream);

/* Tensor-core engine of the same operation (3xTF32, fp32-accurate): requires c1 == 64, c2 == 128, c_in <= 16 and
 * c_out % 128 == 0 (768 / 512 / 256 on the path); other shapes -> VLSAT_ERR_UNSUPPORTED. */
int vlsat_pointnet_tc_fwd(const float* x, int64_t n_obj, int c_in, int64_t n_pts,
                          const float* w1, const float* b1, int c1,
                          const float* w2, const float* b2, int c2,
                          const float* w3, const float* b3, int c_out,
                          float* out, int32_t* argmax, void* stream);

/* A2  Gen_edge_descriptor.message (src/utils/op_utils.py:85-97) with flow='target_to_source':
 *   out[e,0:6] = d[src,0:6]-d[dst,0:6]; out[e,6:11] = log(d[src,6:11]/d[dst,6:11]).  out [E, 11]. */
int vlsat_edge_descriptor_fwd(const float* desc, int64_t n_nodes, const int64_t* edge_index, int64_t n_edges,
                              float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense projections (nn.Linear / Conv1d k=1 everywhere on the path: network_MMG.py:31,59-60,77-79,
 * attention.py:19-22, network_PointNet.py:314-316, SGFN_MMG/model.py:88-111, clip_adapter/model.py:15-17)
 *   y[m, n] = post( act( sum_k x[m,k] w[n,k] + bias[n] + ga[ia[m], n] + gb[ib[m], n] ) )
 *   post(t) = (alpha*t + beta*res[m,n]) * (scale_ptr ? expf(*scale_ptr) : 1)
 * The two row-gather terms implement "project per node, gather per edge" for the concatenations
 * cat[x_i, e, x_j] (network_MMG.py:87) and cat[o[src], o[dst], e] (SGFN_MMG/model.py:260-265).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const float* bias;        /* [N] or NULL                                   */
    const float* gather_a;    /* [*, ld_gather] rows indexed by idx_a, or NULL  */
    const int64_t* idx_a;     /* [M]                                           */
    const float* gather_b;
    const int64_t* idx_b;
    int64_t ld_gather;
    const float* residual;    /* [M, ld_res] or NULL                           */
    int64_t ld_res;
    float alpha;              /* used only when residual != NULL or alpha != 1 */
    float beta;
    const float* scale_ptr;   /* device scalar s: result *= expf(s); or NULL   */
    int act;                  /* VLSAT_ACT_*                                   */
    int bias_per_row;         /* 1: bias is [M] and added per output ROW (used to emit y^T = w x^T) */
    void* split_hi;           /* optional: also emit the (hi, lo) split of the result, compact [M, ld_split], */
    void* split_lo;           /*   so a consuming projection / attention needs no separate split pass         */
    int64_t ld_split;         /*   elements; >= N; multiple of 4 (tf32 pairs) / 8 (bf16 pairs). y may be NULL */
    int split_fmt;            /* VLSAT_SPLIT_TF32: float pairs (vlsat_tf32_split); VLSAT_SPLIT_BF16: bf16 pairs */
} vlsat_epilogue;

/* Operand formats of the tensor-core engines. An fp32 value x travels as the pair (hi, lo):
 *   TF32 pair: hi = x rounded to tf32 (stored as float), lo = x - hi            -> 3 kind::tf32 MMAs, ~2^-22
 *   BF16 pair: hi = bf16(x), lo = bf16(x - hi)  (same 4 bytes per element)     -> 3 kind::f16 MMAs at twice the
 *              tensor rate, |x - hi - lo| <= 2^-17 |x|, fp32 range kept (unlike fp16)                          */
#define VLSAT_SPLIT_TF32 0
#define VLSAT_SPLIT_BF16 1

/* Engine selection. AUTO = tcgen05 BF16x3 whenever the operands are TMA-addressable as bf16 pairs (K % 8 == 0,
 * K >= 32), else tcgen05 3xTF32 (K % 4 == 0, 16-byte aligned rows), else the exact-fp32 FFMA engine (see
 * csrc/gemm_tc.cu). A forced tensor-core engine on an ineligible shape returns VLSAT_ERR_UNSUPPORTED.
 * TC = 3xTF32 (operands are TF32 pairs), TC_BF16X3 (operands are BF16 pairs), TC_1PASS = plain TF32. */
#define VLSAT_ENGINE_AUTO 0
#define VLSAT_ENGINE_SIMT 1
#define VLSAT_ENGINE_TC 2
#define VLSAT_ENGINE_TC_1PASS 3
#define VLSAT_ENGINE_TC_BF16X3 4

typedef struct {
    int engine;               /* VLSAT_ENGINE_*                                                        */
    const void* x_hi;         /* optional pre-split activations, compact [M, K], in the engine's pair format */
    const void* x_lo;         /*   (vlsat_tf32_split / vlsat_bf16_split or an epilogue-emitted split); with  */
    const void* w_hi;         /*   AUTO they are taken to be BF16 pairs when K % 8 == 0, else TF32 pairs      */
    const void* w_lo;         /* optional pre-split weights, compact [N, K] (cache them per parameter)  */
    void* workspace;          /* device scratch for the splits that were not supplied                  */
    size_t workspace_bytes;   /* >= vlsat_linear_workspace_bytes(M, N, K, x_hi == NULL, w_hi == NULL)   */
} vlsat_linear_opts;

/* Process-wide arithmetic of the tensor-core kernels (projections, A9 forward / backward, A8 edge kernel):
 *   VLSAT_PRECISION_FP32 (default): BF16x3 - every product is three bf16 MMAs on (hi, lo) operand pairs, fp32 parity
 *                                   with the reference (north_star: 1e-3 relative);
 *   VLSAT_PRECISION_BF16: one MMA per product on the hi halves - the "bf16" BASELINE configs #3 / #4, which have no
 *                         reference-side counterpart (the reference has no AMP); tolerance against
 *                         the fp32 reference stated in tests/test_bf16_mode_gpu.py. Also set by VLSAT_PRECISION=bf16. */
#define VLSAT_PRECISION_FP32 0
#define VLSAT_PRECISION_BF16 1
int vlsat_set_precision(int mode);
int vlsat_get_precision(void);

int vlsat_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw,
                     float* y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                     const vlsat_epilogue* epi, const vlsat_linear_opts* opts, void* stream);
size_t vlsat_linear_workspace_bytes(int64_t M, int64_t N, int64_t K, int need_x_split, int need_w_split);

/* Backward GEMMs of the dense projections on bf16 (hi, lo) pair operands read AS STORED - no transposed copies
 * (autograd of every nn.Linear / Conv1d(k=1) of the path, e.g. network_MMG.py:87-100; the reference has no
 * hand-written backward). BF16x3 arithmetic, MN-major tcgen05 shared-memory descriptors (csrc/gemm_tc.cu).
 *   VLSAT_GEMM_NN: y [M, N] = a [M, K] . b [K, N]         (dX = dZ W: b is the weight as the forward stores it)
 *   VLSAT_GEMM_TN: y [M, N] = a^T . b, a stored [K, M], b stored [K, N]   (dW = dZ^T X; the reduction over the stored
 *                  rows is split over CTAs when the output has few tiles: deterministic slab sums through `workspace`,
 *                  16-byte aligned, vlsat_gemm_pairs_workspace_bytes bytes)
 * Row strides in elements, multiples of 8; y fp32, 16-byte aligned, ldy % 4 == 0 (overwritten). */
#define VLSAT_GEMM_NN 2
#define VLSAT_GEMM_TN 3
int vlsat_gemm_pairs(int mode, const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo,
                     int64_t ldb, float* y, int64_t ldy, int64_t M, int64_t N, int64_t K, void* workspace,
                     size_t workspace_bytes, void* stream);
size_t vlsat_gemm_pairs_workspace_bytes(int mode, int64_t M, int64_t N, int64_t K);
/* hi = tf32-rounded x (round to nearest), lo = x - hi; x [rows, cols] with row stride ldx, cols % 4 == 0,
 * outputs compact [rows, cols]. */
int vlsat_tf32_split(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* hi, float* lo, void* stream);

/* LayerNorm(x + res) (attention.py:122-123; eps 1e-5), optional ReLU on the way out
 * (network_MMG.py:236-248 fused for the streams whose raw value is not needed). split_hi / split_lo (nullable):
 * also write the bf16 (hi, lo) pair of the result, compact [M, ld_split] (D % 128 == 0 only); y may then be NULL. */
int vlsat_add_layernorm_fwd(const float* x, int64_t ldx, const float* res, int64_t ld_res,
                            const float* gamma, const float* beta, float* y, int64_t ldy,
                            int64_t M, int D, float eps, int relu, void* split_hi, void* split_lo, int64_t ld_split,
                            void* stream);

/* Elementwise helpers: y = relu(x) (network_MMG.py:236-248); row L2 normalisation
 * (SGFN_MMG/model.py:329-330); spatial tail of the 3-D node feature (SGFN_MMG/model.py:296-299):
 * out[n, col0 + 0:6] = desc[n,3:9], out[n, col0+6:8] = log(desc[n,9:11]). */
int vlsat_relu_fwd(const float* x, float* y, int64_t numel, void* stream);
/* y = relu(x) (y nullable) together with the bf16 (hi, lo) pair of the result; numel % 4 == 0. */
int vlsat_relu_pair_fwd(const float* x, float* y, void* split_hi, void* split_lo, int64_t numel, void* stream);
int vlsat_row_l2norm_fwd(const float* x, float* y, int64_t M, int D, void* stream);
int vlsat_spatial_tail_fwd(const float* desc, float* out, int64_t ld_out, int col0, int64_t n_nodes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * A6 + A7  node attention inside scenes (network_MMG.py:181-205,217-218; attention.py:54-77).
 *   seg_start/seg_end [n] int32: scene range of every node, derived from batch_ids (non-decreasing).
 *   out[a, h*dk:(h+1)*dk] = sum_b softmax_b( q_a.k_b/sqrt(dk) + MLP([c_b-c_a, |c_b-c_a|])[h] ) v_b,
 *   b over the scene of a. MLP = Linear(4,32) ReLU LN Linear(32,32) ReLU LN Linear(32,H)
 *   (network_MMG.py:165-173); fc_w packs its 10 tensors (layout in DESIGN.md / ops.py).
 *   err_flag (device int32): set to 1 if batch_ids is not non-decreasing.
 * ---------------------------------------------------------------------------------------------- */
int vlsat_scene_ranges(const int64_t* batch_ids, int64_t n_nodes, int32_t* seg_start, int32_t* seg_end,
                       int32_t* err_flag, void* stream);
int vlsat_node_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                        const float* centres, int64_t ld_centres,
                        const int32_t* seg_start, const int32_t* seg_end,
                        const float* fc_w, int n_heads, int dk,
                        float* out, int64_t ldo, int64_t n_nodes, int skip_scenes_upto, void* stream);
/* Scene-resident form of the same attention for scenes of up to vlsat_node_bias_table_max_scene() (= 64) nodes.
 * The bias depends only on the centres and MMG.forward evaluates 2*depth attentions over the same scenes, so it is
 * computed once: vlsat_node_bias_table fills table[(h*n_nodes + a)*64 + j] = MLP(...)[h] for query a and the j-th node of its
 * scene (rows of larger scenes are left untouched); vlsat_node_attn_scene_fwd then serves every scene of <= 64 nodes
 * with one CTA per (scene, head) holding that head's Q, K, V rows in shared memory, and leaves the rows of larger
 * scenes untouched: call vlsat_node_attn_fwd(..., skip_scenes_upto = 64) for those (it skips the scenes already
 * served; with skip_scenes_upto = 0 it serves everything). table: n_nodes * 64 * H floats. */
int vlsat_node_bias_table_max_scene(void);
int vlsat_node_bias_table(const float* centres, int64_t ld_centres, const int32_t* seg_start, const int32_t* seg_end,
                          const float* fc_w, int n_heads, float* table, int64_t n_nodes, void* stream);
int vlsat_node_attn_scene_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                              const float* table, const int32_t* seg_start, const int32_t* seg_end, int n_heads, int dk,
                              float* out, int64_t ldo, int64_t n_nodes, void* stream);

/* A7 with the reference's DENSE arguments (ScaledDotProductAttention.forward, attention.py:41-78, as MMG.forward calls it,
 * network_MMG.py:217-218): out = softmax(mask(q k^T / sqrt(dk) (*|+) weights)) v.
 *   weights [H, nq, nk] fp32 (head stride given; way: 0 none, 1 'mul', 2 'add'), mask fp32, 0 = masked, [nq, nk] shared
 *   by all heads (mask_head_stride 0) or per head; a fully masked row yields NaN like torch.softmax does.
 * Exact fp32; exists for module-level drop-in of MultiHeadAttention under the reference's own MMG - the fast path
 * (vlsat_node_attn_scene_fwd) never builds these tensors. */
int vlsat_dense_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                         const float* weights, int64_t weights_head_stride, int way, const float* mask,
                         int64_t mask_head_stride, float* out, int64_t ldo, int64_t nq, int64_t nk, int n_heads,
                         int dk, void* stream);

/* A9  cross_attn_rel core (network_MMG.py:231; attention.py:54-77 without mask/bias): streaming
 * softmax(QK^T/sqrt(dk)) V over ALL keys, never materialising the [H, nq, nk] score tensor.
 * lse (nullable) [H, nq] = log-sum-exp per row (saved for backward). */
int vlsat_flash_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                         float* out, int64_t ldo, float* lse, int64_t nq, int64_t nk, int n_heads, int dk,
                         void* stream);

/* Tensor-core engine of the same operation (3xTF32, fp32-accurate; dk must be 64): operands are the tf32
 * hi/lo splits (vlsat_tf32_split) of Q [nq, H*64], K [nk, H*64] and of the TRANSPOSED value projection
 * V^T [H*64, nk] (row stride ldvt, a multiple of 4), which the dense projection emits directly by swapping
 * its operands (y^T = W_v x^T with bias_per_row). */
int vlsat_flash_attn_tc_fwd(const float* q_hi, const float* q_lo, int64_t ldq,
                            const float* k_hi, const float* k_lo, int64_t ldk,
                            const float* vt_hi, const float* vt_lo, int64_t ldvt,
                            float* out, int64_t ldo, float* lse, int64_t nq, int64_t nk, int n_heads, int dk,
                            void* stream);

/* BF16x3 engine of the same operation (default for dk = 64): operands are bf16 (hi, lo) pairs (vlsat_bf16_split:
 * hi = bf16(x), lo = bf16(x - hi); uint16 storage) of Q [nq, H*64], K [nk, H*64], V^T [H*64, nk]; row strides in
 * elements, multiples of 8. Halves the L2->SM traffic that bounds this kernel and doubles the MMA rate relative
 * to the 3xTF32 form; product error 2^-17 relative (see csrc/flash_attn_bf16.cu). */
int vlsat_bf16_split(const float* x, int64_t ldx, int64_t rows, int64_t cols, void* hi, void* lo, int64_t ld_out,
                     void* stream);
int vlsat_flash_attn_bf16x3_fwd(const void* q_hi, const void* q_lo, int64_t ldq,
                                const void* k_hi, const void* k_lo, int64_t ldk,
                                const void* vt_hi, const void* vt_lo, int64_t ldvt,
                                float* out, int64_t ldo, float* lse, int64_t nq, int64_t nk, int n_heads, int dk,
                                void* workspace, size_t workspace_bytes, void* stream);
/* The kernel splits the key range over CTAs when that fills the 148 SMs in fewer rounds; the partial softmax
 * states then go through `workspace` (16-byte aligned, size below; 0 = no split needed). */
size_t vlsat_flash_attn_bf16x3_workspace_bytes(int64_t nq, int64_t nk, int n_heads);

/* A9 backward, streaming (round 2): gradients of softmax(QK^T/sqrt(dk)) V w.r.t. Q, K, V - autograd through
 * attention.py:41-78 as called at network_MMG.py:231 (no mask, no bias). Scores and probabilities never leave the
 * SM. Two launches of one tcgen05 kernel (csrc/flash_attn_bwd.cu): key tiles stationary -> dK, dV; query tiles
 * stationary -> dQ. BF16x3 arithmetic like the forward.
 *   vlsat_bf16_split_t       fp32 [rows, cols] -> bf16 (hi, lo) pairs [rows, cols] (row stride ld_out) and, when hi_t / lo_t
 *                            are given, the transposed pairs [cols, ld_t] (ld_t >= rows, multiple of 8, tail zero)
 *   vlsat_flash_attn_bwd_stats  lse2 = lse*log2(e), delta = rowsum_head(dout o out); both [H, ld_stat], ld_stat a
 *                            multiple of 64 >= nq, padding written as (1e30, 0)
 *   vlsat_flash_attn_bf16x3_bwd  dq [nq, H*64], dk / dv [nk, H*64] fp32 (overwritten); head_dim must be 64;
 *                            workspace (16-byte aligned) of vlsat_flash_attn_bf16x3_bwd_workspace_bytes bytes when the
 *                            streamed range is split over CTAs to fill the 148 SMs (deterministic slab sums). */
typedef struct vlsat_bf16_pair { const void* hi; const void* lo; int64_t ld; } vlsat_bf16_pair;
typedef struct vlsat_flash_bwd_operands {
    vlsat_bf16_pair q, k, v, dout;          /* [n, H*64] */
    vlsat_bf16_pair q_t, k_t, dout_t;       /* [H*64, n] transposed copies */
    const float* lse2; const float* delta; int64_t ld_stat;
} vlsat_flash_bwd_operands;
int vlsat_bf16_split_t(const float* x, int64_t ldx, int64_t rows, int64_t cols, void* hi, void* lo, int64_t ld_out,
                       void* hi_t, void* lo_t, int64_t ld_t, void* stream);
int vlsat_flash_attn_bwd_stats(const float* dout, int64_t ld_dout, const float* out, int64_t ld_out, const float* lse,
                               int64_t ld_lse, float* lse2, float* delta, int64_t ld_stat, int64_t nq, int n_heads, int dk,
                               void* stream);
int vlsat_flash_attn_bf16x3_bwd(const vlsat_flash_bwd_operands* operands, float* dq, int64_t ld_dq, float* dk, int64_t ld_dk,
                                float* dv, int64_t ld_dv, int64_t nq, int64_t nk, int n_heads, int head_dim,
                                void* workspace, size_t workspace_bytes, void* stream);
size_t vlsat_flash_attn_bf16x3_bwd_workspace_bytes(int64_t nq, int64_t nk, int n_heads);

/* ------------------------------------------------------------------------------------------------
 * A8  graph attention layer core (network_MMG.py:34-41,96-104; network_util.py:50-73).
 *   vlsat_build_csr: stable counting sort of edges by index_row (= edge_index[0] for
 *     flow='target_to_source'): row_ptr [n_nodes+1], perm [n_edges] (edge ids grouped by node,
 *     ascending inside a group). workspace >= (n_nodes+1+n_edges)*4 bytes. Bit-exact integer work.
 *   vlsat_gat_edge_fwd: for every edge e (src,dst) and head h
 *       u = [q[src, c*H+h]]_c ++ [k[e, c*H+h]]_c ;  t = C2 . ReLU(C1 . u + c1) + c2 ; p = softmax(t)
 *       m[e, c*H+h] = p[c] * v[dst, c*H+h]
 *     xx[n,:] = aggr_{e: src(e)=n} m[e,:]  (max: 0 for nodes without out-edges; add; mean).
 *     q = proj_query(x) [n_nodes, H*d_n], v = proj_value(x) [n_nodes, H*d_o], k = proj_edge(e)
 *     [n_edges, H*d_e] (use_edge=0: k ignored, C1 is [hid, d_n]); ldq/ldv/ldk/ld_xx = row strides, so q, v
 *     may be column slices of one fused node projection and xx a slice of the cat[x, xx] buffer.
 *     prob (nullable) [n_edges, d_o, H]; argmax (nullable) [n_nodes, H*d_o] int32 edge id or -1.
 * ---------------------------------------------------------------------------------------------- */
int vlsat_build_csr(const int64_t* index_row, int64_t n_edges, int64_t n_nodes,
                    int32_t* row_ptr, int32_t* perm, void* workspace, size_t workspace_bytes, void* stream);
int vlsat_gat_edge_fwd(const float* q, int64_t ldq, const float* v, int64_t ldv, const float* k, int64_t ldk,
                       const int64_t* edge_index, const int32_t* row_ptr, const int32_t* perm,
                       const float* c1, const float* c1_bias, const float* c2, const float* c2_bias,
                       int64_t n_nodes, int64_t n_edges, int n_heads, int d_n, int d_e, int d_o, int hid,
                       int aggr, int use_edge, float* xx, int64_t ld_xx, float* prob, int32_t* argmax, void* stream);

/* Tensor-core engine of vlsat_gat_edge_fwd for aggr = max with edge features (the mmgnet.json layer), BF16x3
 * (csrc/gat_tc.cu). Edges must be in source-sorted (CSR) order: src_sorted[i] non-decreasing; every per-edge
 * operand uses that order. Operands are "head-major" (row (e,h) or (n,h) contiguous), prepared by projections
 * with permuted weight rows (see cvpr2023-vlsat_b200/gat.py):
 *   k_hi/k_lo [E*H, d_e] bf16 pair of proj_edge(e) (as emitted by vlsat_linear_fwd with VLSAT_SPLIT_BF16);
 *   qc: row (n,h) = qc + n*ld_qc + h*hid holds C1[:, :d_n] . q3[n,:,h] + c1_bias;  v: row (n,h) = v + n*ld_v + h*d_o;
 *   c1k = C1[:, d_n:] [hid, d_e] and c2 [d_o, hid] as bf16 pairs (vlsat_bf16_split).
 *   xx [n_nodes, H*d_o] (row stride ld_xx) is written in the reference's interleaved order c*H + h; nodes without
 *   outgoing edges get 0.  prob (nullable) [E, d_o, H] in sorted order.
 * workspace >= n_nodes * H * d_o * 4 bytes; workspace_ready = 1 promises that every word of it holds INT_MIN (the
 *   state this call leaves it in), 0 makes the call fill it first.
 * Constraints: 128 % H == 0, d_e == 64, hid % 32 == 0, hid <= 128, d_o in {32, 64};
 * other shapes -> VLSAT_ERR_UNSUPPORTED (use vlsat_gat_edge_fwd). */
int vlsat_gat_edge_tc_fwd(const void* k_hi, const void* k_lo, const float* qc, int64_t ld_qc,
                          const float* v, int64_t ld_v, const int64_t* src_sorted, const int64_t* dst_sorted,
                          const void* c1k_hi, const void* c1k_lo, const void* c2_hi, const void* c2_lo,
                          const float* c2_bias, int64_t n_nodes, int64_t n_edges, int n_heads, int d_e, int hid,
                          int d_o, float* xx, int64_t ld_xx, float* prob, void* workspace, size_t workspace_bytes,
                          int workspace_ready, void* stream);

/* Edge-order bookkeeping (bit-exact): out[i,:] = in[idx[i],:] (gather=1) or out[idx[i],:] = in[i,:] (gather=0)
 * for fp32 rows (cols % 4 == 0), and out[:, i] = edge_index[:, perm[i]] for the int64 [2, E] edge list.
 * split_hi / split_lo (nullable): also write the bf16 (hi, lo) pair of the permuted rows, compact [rows, cols]. */
int vlsat_permute_rows(const float* in, int64_t ld_in, const int32_t* idx, int64_t rows, int cols,
                       float* out, int64_t ld_out, int gather, void* split_hi, void* split_lo, void* stream);
int vlsat_permute_edges(const int64_t* edge_index, const int32_t* perm, int64_t n_edges, int64_t* out, void* stream);

/* ================================================================================================
 * Backward / training-mode entry points. The reference has no hand-written backward: torch.autograd
 * over the modules cited above defines it (training step: SGFN_MMG/model.py:337-346,483-488). The
 * GEMM-shaped parts of every backward (dX = dZ W, dW = dZ^T X) are vlsat_linear_fwd calls on operands
 * transposed by vlsat_transpose; the entry points below are the remaining memory-bound pieces.
 * Buffers documented as "accumulated" must be zero-filled by the caller.
 * ================================================================================================ */

/* out[b, c, r] = in[b, r, c] for `batch` matrices [rows, cols]; ld_out >= rows and the tail columns
 * [rows, ld_out) of every output row are zero-filled (GEMM reduction lengths are multiples of 4). */
int vlsat_transpose(const float* in, int64_t ld_in, int64_t batch_stride_in, float* out, int64_t ld_out,
                    int64_t batch_stride_out, int64_t batch, int64_t rows, int64_t cols, void* stream);

/* Backward of the projection epilogue act(.)*scale*exp(*scale_ptr): dz = dy * act'(y) * scale * exp(*scale_ptr)
 * (y = forward OUTPUT of the activation; dz nullable, may alias dy) and dbias[n] += sum_m dz[m, n]
 * (nullable, accumulated). nn.Linear + ReLU / sigmoid sites: network_PointNet.py:328-341, network_MMG.py:31-32,59-60. */
int vlsat_act_bwd(const float* dy, int64_t lddy, const float* y, int64_t ldy, int act, float scale,
                  const float* scale_ptr, float* dz, int64_t lddz, float* dbias, int64_t M, int64_t N, void* stream);
/* Same, four columns per thread (N % 4 == 0, 16-byte aligned rows), optionally writing the bf16 (hi, lo) pair of dz in the
 * same pass (row stride ld_split): the operand format of the two backward GEMMs that read dz next (vlsat_gemm_pairs).
 * dz, dbias and the pair are each nullable (at least one must be given). */
int vlsat_act_bwd_pair(const float* dy, int64_t lddy, const float* y, int64_t ldy, int act, float scale,
                       const float* scale_ptr, float* dz, int64_t lddz, float* dbias, void* split_hi, void* split_lo,
                       int64_t ld_split, int64_t M, int64_t N, void* stream);

/* Weight gradient of a small projection over a tall operand pair: dw[n, k] += sum_m dz[m, n] * x[m, k] (accumulated),
 * N, K <= 128 (the per-(edge, head) attention MLP network_MMG.py:73 and the PointNet convs network_PointNet.py:99-100,
 * whose [N, K] result would be a single tensor-core tile). Exact FP32. */
int vlsat_wgrad_small(const float* dz, int64_t lddz, const float* x, int64_t ldx, int64_t M, int N, int K, float* dw,
                      int64_t lddw, void* stream);

/* Row gather and its backward. Row i reads / accumulates into row idx[i / R] * R + i % R (R = rows_per_idx;
 * R = H addresses rows (node, head) of a head-major tensor from rows (edge, head)). Backward of
 * Gen_Index (network_util.py:50-62) and of the "project per node, gather per edge" epilogue. out accumulated. */
int vlsat_gather_rows(const float* in, int64_t ld_in, const int64_t* idx, int rows_per_idx, int64_t rows, int cols,
                      float* out, int64_t ld_out, void* stream);
int vlsat_scatter_add_rows(const float* in, int64_t ld_in, const int64_t* idx, int rows_per_idx, int64_t rows, int cols,
                           float* out, int64_t ld_out, void* stream);

/* Backward of vlsat_add_layernorm_fwd (attention.py:122-123): dx = d(x + res); dgamma / dbeta accumulated. */
int vlsat_add_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* res, int64_t ld_res,
                            const float* gamma, const float* beta, float* dx, int64_t lddx, float* dgamma, float* dbeta,
                            int64_t M, int D, float eps, int relu, void* stream);

/* A8, differentiable decomposition (network_MMG.py:96-104 + Aggre_Index network_util.py:64-73). Edges in CSR order.
 *   t [E*H, d_o]: attention-MLP output of row (e, h);  v: row (n, h) = v + n*ldv + h*d_o (head-major proj_value)
 *   prob[(e,h), :] = softmax(t[(e,h), :]);  m[e, c*H+h] = prob[(e,h), c] * v[dst(e), h, c];  xx[n, c*H+h] = aggr m
 *   argmax [n_nodes, H*d_o] int32 (max only): winning edge or -1.
 * backward: dt [E*H, d_o] written; dv [n_nodes, ld_dv] accumulated. */
int vlsat_gat_softmax_aggr_fwd(const float* t, const float* v, int64_t ldv, const int64_t* dst_sorted,
                               const int32_t* row_ptr, int64_t n_nodes, int64_t n_edges, int n_heads, int d_o,
                               int aggr, float* xx, int64_t ld_xx, float* prob, int32_t* argmax, void* stream);
int vlsat_gat_softmax_aggr_bwd(const float* dxx, int64_t ld_dxx, const float* prob, const float* v, int64_t ldv,
                               const int64_t* dst_sorted, const int32_t* row_ptr, const int32_t* argmax,
                               int64_t n_nodes, int64_t n_edges, int n_heads, int d_o, int aggr, float* dt,
                               float* dv, int64_t ld_dv, void* stream);

/* A9 backward, score stage for one head and one block of queries (attention.py:54-77): s = Q K^T and dp = dO V^T
 * are [nq, nk] projection outputs; P = exp(scale*s - lse), dS = P*(dp - delta)*scale. Writes ds [nq, nk] (may alias
 * dp), ds_t = dS^T and p_t = P^T [nk, ld_t] (columns >= nq zero-filled). delta[i] = dO_i . O_i (vlsat_rowdot_heads,
 * out [H, M]). */
int vlsat_attn_prob_bwd(const float* s, const float* dp, int64_t ld, const float* lse, const float* delta, float scale,
                        float* ds, float* ds_t, float* p_t, int64_t ld_t, int64_t nq, int64_t nk, void* stream);
/* Same stage with dS, dS^T and P^T written as the bf16 (hi, lo) pairs the three following products consume
 * (ds_* [nq, ld_ds], dst_* / pt_* [nk, ld_t], columns >= nq zero-filled); no fp32 [nq, nk] output. */
int vlsat_attn_prob_bwd_pairs(const float* s, const float* dp, int64_t ld, const float* lse, const float* delta, float scale,
                              void* ds_hi, void* ds_lo, int64_t ld_ds, void* dst_hi, void* dst_lo, void* pt_hi, void* pt_lo,
                              int64_t ld_t, int64_t nq, int64_t nk, void* stream);
int vlsat_rowdot_heads(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t M, int n_heads,
                       int dk, void* stream);

/* A6 + A7 with the distance bias as an explicit per-pair tensor (differentiable form of vlsat_node_attn_fwd):
 * bias row of (query a, key b) = pair_off[a] + (b - seg_start[a]), [*, H]. vlsat_pair_features writes the MLP input
 * [c_b - c_a, |c_b - c_a|] per pair (network_MMG.py:189-196). max_scene >= the largest scene (host-known bound).
 * backward: dq written; dk_out, dv_out accumulated; dbias [pairs, H] written. */
int vlsat_pair_features(const float* centres, int64_t ld_centres, const int32_t* seg_start, const int32_t* seg_end,
                        const int64_t* pair_off, int64_t n_nodes, float* out, void* stream);
int vlsat_node_attn_bias_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                             const float* bias, const int64_t* pair_off, const int32_t* seg_start,
                             const int32_t* seg_end, int n_heads, int dk, int max_scene, float* out, int64_t ldo,
                             int64_t n_nodes, void* stream);
int vlsat_node_attn_bias_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                             const float* bias, const int64_t* pair_off, const int32_t* seg_start,
                             const int32_t* seg_end, const float* dout, int64_t lddo, int n_heads, int dk,
                             int max_scene, float* dq, int64_t lddq, float* dk_out, int64_t lddk, float* dv_out,
                             int64_t lddv, float* dbias, int64_t n_nodes, void* stream);

/* A1 backward through the max-pool (network_PointNet.py:164): dz3 [n_obj, c_out] = dOut masked by out > 0,
 * argmax from the forward, h2 [n_obj*n_pts, c2] = post-ReLU layer-2 activations (recomputed by two projections).
 * dw3 [c_out, c2] and dh2 [n_obj*n_pts, c2] accumulated. */
int vlsat_pointnet_pool_bwd(const float* dz3, const int32_t* argmax, const float* h2, const float* w3, int64_t n_obj,
                            int64_t n_pts, int c_out, int c2, float* dw3, float* dh2, void* stream);

/* nn.Dropout in training mode (8 sites on the path). Counter-based mask: element i is kept iff
 * hash(seed, *device_step, offset + i) >= p * 2^32; kept values are scaled by 1/(1-p). The backward is the same call
 * on dy. device_step (nullable): uint64 counter in device memory, read at run time - a replayed CUDA graph of the
 * training step bumps it so that every replay draws fresh masks. */
int vlsat_dropout(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t rows, int64_t cols, float p,
                  uint64_t seed, uint64_t offset, const uint64_t* device_step, void* stream);
/* The same pass also writing the bf16 (hi, lo) pair of y, compact rows of ld_pair elements (cols % 4 == 0, 16-byte aligned x / y
 * rows, ld_pair % 4 == 0): the operand a following projection reads, without a separate split pass. */
int vlsat_dropout_pair(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t rows, int64_t cols, float p,
                       uint64_t seed, uint64_t offset, const uint64_t* device_step, void* pair_hi, void* pair_lo,
                       int64_t ld_pair, void* stream);

/* nn.BatchNorm1d of mlp_3d (SGFN_MMG/model.py:108). batch_stats = 1: mean / rstd computed from x (biased variance) and
 * written, running stats (nullable) updated with `momentum` and the unbiased variance; batch_stats = 0: mean / rstd
 * are inputs (running stats). Optional fused ReLU. backward: dgamma / dbeta written. */
int vlsat_batchnorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float* mean, float* rstd,
                        float* running_mean, float* running_var, float momentum, float eps, int batch_stats, int relu,
                        float* y, int64_t ldy, int64_t M, int64_t N, void* stream);
int vlsat_batchnorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                        const float* gamma, const float* beta, int relu, int batch_stats, float* dx, int64_t lddx,
                        float* dgamma, float* dbeta, int64_t M, int64_t N, void* stream);

/* Backward of vlsat_row_l2norm_fwd, and out[0] += a . b (gradient of obj_logit_scale, SGFN_MMG/model.py:327-330). */
int vlsat_row_l2norm_bwd(const float* dy, const float* x, float* dx, int64_t M, int D, void* stream);
int vlsat_dot_accum(const float* a, const float* b, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N1  training-step glue: the losses of Mmgnet.process_train (src/model/SGFN_MMG/model.py:343-418)
 * and the optimiser step of Mmgnet.backward (:483-488; groups and schedule :143-158).
 * Every *_fwd writes ONE loss value per row into row_loss; vlsat_sum_rows adds a buffer up in a fixed order
 * (reproducible bit for bit) and applies the mean / term coefficient. Every *_bwd multiplies by gout[0] (the
 * upstream gradient of the scalar loss, read from DEVICE memory: no host sync) and by `coef` (term weight /
 * element count of the mean), and WRITES (does not accumulate) the gradient.
 * ---------------------------------------------------------------------------------------------- */
/* F.cross_entropy(logits [R, C], target int64 [R]) (:343-344): row_loss = logsumexp(x) - x[target], row_lse kept for
 * the backward: dlogits = gout * coef * (softmax(x) - onehot). Rows whose target is outside [0, C) give 0 / 0. */
int vlsat_cross_entropy_fwd(const float* logits, int64_t ld, const int64_t* target, int64_t R, int C,
                            float* row_loss, float* row_lse, void* stream);
int vlsat_cross_entropy_bwd(const float* logits, int64_t ld, const int64_t* target, const float* row_lse,
                            const float* gout, float coef, float* dlogits, int64_t ldd, int64_t R, int C, void* stream);
/* WEIGHT_EDGE == 'DYNAMIC' (:353-366): weight[c] = |scale / (log(sum_e gt[e, c] + 1) + 1)|, gt [E, C] of 0/1 floats,
 * C <= 64; scale = 1 (1e-2 with ignore_none_rel). */
int vlsat_rel_class_weights(const float* gt, int64_t E, int C, float scale, float* weight, void* stream);
/* F.binary_cross_entropy(p [E, C], y [E, C], weight [C] or NULL) (:375-376), logs clamped at -100 like ATen;
 * backward dp = gout * coef * w (p - y) / max((1 - p) p, 1e-12). */
int vlsat_bce_fwd(const float* p, const float* y, const float* weight, int64_t E, int C, float* row_loss, void* stream);
int vlsat_bce_bwd(const float* p, const float* y, const float* weight, const float* gout, float coef, float* dp,
                  int64_t E, int C, void* stream);
/* cosine_loss (:257-258) after the row normalisations of :402-403: row_loss = max(margin - cos(a_r, b_r), 0).
 * da / db nullable. */
int vlsat_cosine_margin_fwd(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t R, int D, float margin,
                            float* row_loss, void* stream);
int vlsat_cosine_margin_bwd(const float* a, int64_t lda, const float* b, int64_t ldb, const float* gout, float coef,
                            float margin, float* da, int64_t ldda, float* db, int64_t lddb, int64_t R, int D, void* stream);
/* F.l1_loss(x / |x|, target) (:409-410): row_loss = sum_c |x_c / |x| - t_c|. */
int vlsat_l1_unit_fwd(const float* x, int64_t ldx, const float* target, int64_t ldt, int64_t R, int D, float* row_loss, void* stream);
int vlsat_l1_unit_bwd(const float* x, int64_t ldx, const float* target, int64_t ldt, const float* gout, float coef,
                      float* dx, int64_t lddx, int64_t R, int D, void* stream);
/* term = scale * sum_i v[i] (single CTA, fixed order); out_term[0] = term (nullable);
 * out_total[0] = (accumulate ? out_total[0] : 0) + total_coef * term (nullable) - the weighted sum of :412. */
int vlsat_sum_rows(const float* v, int64_t n, float scale, float* out_term, float* out_total, float total_coef,
                   int accumulate, void* stream);

/* One tensor of the multi-tensor AdamW step; the table lives in device memory. vmax: amsgrad state or NULL. */
typedef struct {
    float* p; const float* g; float* m; float* v; float* vmax;
    int64_t n;
    float lr;               /* base learning rate of the tensor's parameter group (:143-156) */
    float weight_decay;
} vlsat_adamw_tensor;
/* torch.optim.AdamW.step() for every tensor of the table + CosineAnnealingLR(T_max = t_max, eta_min = 0) (:157, :483-488):
 * with k = step[0] + 1, lr_k = lr * (1 + cos(pi (k - 1) / t_max)) / 2 (t_max <= 0: constant), p *= 1 - lr_k wd,
 * m += (1 - beta1)(g - m), v = beta2 v + (1 - beta2) g^2, p -= lr_k / (1 - beta1^k) * m / (sqrt(v) / sqrt(1 - beta2^k) + eps);
 * then step[0] = k. Chunk c covers elements [chunk_index[c] * chunk_elems, +chunk_elems) of tensor chunk_tensor[c]
 * (one CTA each). The step counter lives in device memory, so the call can be replayed inside a CUDA graph. */
int vlsat_adamw_step(const vlsat_adamw_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                     int64_t n_chunks, int chunk_elems, double beta1, double beta2, float eps, int64_t* step,
                     int64_t t_max, void* stream);

/* N4: text supervision target of get_rel_emb (SGFN_MMG/model.py:221-255) from a CACHED table of prompt features
 * instead of a python loop + CLIP text encoder per step. table [S, O, R + 1, dim] fp32: features of "a point cloud of a {s}
 * {rel} a {o}" at [s, o, r] and of "the {s} and the {o} has no relation in the point cloud" at [s, o, R] (filled once by the
 * caller's text encoder). out[e] = normalise(mean over the ground-truth predicates of edge e), or the no-relation row;
 * edges [E, 2] int64 (subject, object), gt_rel [E, R] 0/1. dim in {256, 512, 768}. */
int vlsat_rel_text_embed(const float* table, int n_obj_cls, int n_rel_cls, int dim, const int64_t* gt_cls,
                         const float* gt_rel, int64_t ld_rel, const int64_t* edges, int64_t E, float* out,
                         int64_t ld_out, void* stream);

/* Data-parallel training (the reference is single-process, SGFN_MMG/model.py:483-488 runs backward() and
 * optimizer.step() back to back): dst_t = scale * src_t for every tensor of a table in ONE launch - packs a step's
 * gradients into the flat buffer the NCCL all-reduce runs on, with 1 / world_size folded in. */
typedef struct { float* dst; const float* src; int64_t n; } vlsat_copy_tensor;
int vlsat_pack_scale(const vlsat_copy_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                     int64_t n_chunks, int chunk_elems, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N2  device-side object preparation of the input pipeline: for object o and sampled point p,
 *   row = cloud[choice[o, p], :]   (src/dataset/dataset_3dssg.py:288-289; `choice` = the np.random.choice result offset
 *                                   into the scan cloud, int64 [n_obj, n_pts]; out-of-range indices are clamped)
 *   descriptor[o] = [mean xyz, unbiased std xyz, max - min xyz, product of the extents, largest extent]
 *                                   (src/utils/op_utils.py:47-64 gen_descriptor, computed BEFORE centring as :291 does)
 *   obj_points[o, c, p] = row[c] - (c < 3 ? mean[c] : 0)      (zero_mean :293 + permute(0, 2, 1) of src/model/model.py:71)
 * cloud [n_cloud, n_channels] fp32 (xyz first; rgb / normals after), obj_points [n_obj, n_channels, n_pts], descriptor [n_obj, 11].
 * ---------------------------------------------------------------------------------------------- */
int vlsat_object_prep_fwd(const float* cloud, int64_t ld_cloud, int64_t n_cloud, int n_channels, const int64_t* choice,
                          int64_t n_obj, int64_t n_pts, float* obj_points, float* descriptor, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N3  evaluation ranks of --mode eval (src/utils/eva_utils_acc.py; Mmgnet.process_val, SGFN_MMG/model.py:463-472).
 * rank = 1 + #{scores strictly greater than the ground truth's}, capped at topk + 1 - counted, never sorted.
 * First run on a B200 in round 2 (tests/test_eval_ranks_gpu.py, bit-exact against the oracle and the reference's fixtures).
 * ---------------------------------------------------------------------------------------------- */
/* y = softmax(x) over each row (F.softmax(objs_pred, dim=-1), eva_utils_acc.py:143-145); y compact [R, C]. */
int vlsat_softmax_rows(const float* x, int64_t ld, int64_t R, int C, float* y, void* stream);
/* evaluate_topk_object (:27-39): ranks[n] int32. */
int vlsat_topk_object_ranks(const float* pred, int64_t ld, const int64_t* target, int64_t N, int C, int topk,
                            int32_t* ranks, void* stream);
/* evaluate_topk_predicate (:42-79) on get_gt's multi-label targets (:6-24): rel_prob, gt_rel [E, C] (C <= 64); ranks [E, C]
 * int32: row e holds the edge's entries (one per ground-truth label, or one if it has none: first position below
 * `threshold`) sorted ascending with the reference's "i-th smallest minus i" adjustment (:73-78; tied labels can take it
 * to 0 and below), INT32_MIN beyond. */
int vlsat_topk_predicate_ranks(const float* rel_prob, const float* gt_rel, int64_t E, int C, int topk, float threshold,
                               int32_t* ranks, void* stream);
/* evaluate_triplet_topk (:137-211), ranks only: score(i, j, k) = (obj_prob[sub, i] * obj_prob[obj, j]) * rel_prob[e, k] in
 * fp32 round-to-nearest (the reference's association order); edges [E, 2] int64 (subject, object) as process_val passes
 * them; obj_prob [n_nodes, n_obj_cls] already softmaxed; ranks [E, n_rel_cls] in the layout of vlsat_topk_predicate_ranks. */
int vlsat_topk_triplet_ranks(const float* obj_prob, int64_t n_nodes, int n_obj_cls, const float* rel_prob, int n_rel_cls,
                             const int64_t* gt_cls, const float* gt_rel, const int64_t* edges, int64_t E, int topk,
                             float threshold, int32_t* ranks, void* stream);

/* 100 * #{ranks <= t_j} / #{ranks} for three thresholds over a rank array in either layout above (INT32_MIN = empty slot):
 * the Obj_R1/5/10 and Pred_R1/3/5 figures process_train logs every step (SGFN_MMG/model.py:422-432), without a host
 * sync. out3 [3] float; one CTA, deterministic. */
int vlsat_recall_at(const int32_t* ranks, int64_t n, int t0, int t1, int t2, float* out3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLSAT_B200_H */
