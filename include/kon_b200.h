/*
 * kon_b200.h -- C-ABI of libkon_b200.so: hand-written sm_100a kernels for the CTR
 * hot path of TIXhjq/ML_Function (`kon.model.ctr_model`).
 *
 * The reference has no native code and no FFI: its hot path is Keras layers whose
 * arithmetic runs inside TensorFlow 2.1.  Each entry point below replaces the TF op
 * stream one reference `call` dispatches (cited as IL/BL/CL = interactive_layer.py /
 * behavior_layer.py / core_layer.py under kon/model/ctr_model/layer/...), and is what
 * a maintainer would bind from Python with ctypes.
 *
 * Conventions
 *   - Tensors cross the boundary as `const DLTensor*` (DLPack ABI struct: plain
 *     pointers, sizes and strides -- no framework types).  The caller owns every
 *     buffer, output and workspace; the library never allocates device memory,
 *     never keeps a pointer past the call, never calls a DLPack deleter and never
 *     synchronises the device.  All work is enqueued on `stream` (a cudaStream_t).
 *   - Return value: 0 on success, negative KON_E* on failure; the message is in the
 *     thread-local `kon_last_error()`.  Nothing throws across the ABI.
 *   - Shapes/dtypes/devices/alignment are validated on the host before any launch.
 *   - There is no CPU implementation behind any of these symbols.
 */
#ifndef KON_B200_H_
#define KON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- DLPack tensor (ABI-identical to dlpack.h's DLTensor; defined here so that the
 *      header is self-contained; skipped when dlpack.h was included first) ---------- */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0, kDLUInt = 1, kDLFloat = 2, kDLBfloat = 4 } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void*      data;
  DLDevice   device;
  int32_t    ndim;
  DLDataType dtype;
  int64_t*   shape;
  int64_t*   strides;      /* in elements; NULL = compact row-major */
  uint64_t   byte_offset;
} DLTensor;
#endif

#define KON_ABI_VERSION 1

enum {
  KON_OK = 0,
  KON_EINVAL = -1,       /* bad shape / dtype / stride / alignment / null pointer */
  KON_EDEVICE = -2,      /* tensor not on a CUDA device, or devices differ */
  KON_EUNSUPPORTED = -3, /* legal request outside what the kernels cover */
  KON_ECUDA = -4,        /* CUDA runtime error at launch */
  KON_EWORKSPACE = -5    /* workspace too small */
};

int         kon_abi_version(void);
const char* kon_last_error(void);
/* Number of kernels of THIS library launched by the process so far (every __global__ of
 * libkon_b200 counts once per launch). */
long long   kon_launch_count(void);
/* Per-kernel device timing for roofline reports.  While enabled, entry points that launch
 * several kernels (kon_cin_fwd/bwd, kon_embed_fwd/bwd) bracket their main kernels with CUDA
 * events recorded on the caller's stream (not usable under stream capture).  kon_profile_read
 * synchronises on the recorded events and returns the summed duration and the launch count of
 * one kernel name ("cin_fwd_tc_kernel", "cin_dw_tc_kernel", "cin_da_tc_kernel",
 * "embed_fwd_vec_kernel", "embed_bwd_sort", "embed_reduce_kernel"). */
int         kon_profile_enable(int on);
int         kon_profile_reset(void);
int         kon_profile_read(const char* kernel, double* total_ms, long long* launches);
/* SM count / arch of the device the tensors live on (for host-side grid sizing). */
int         kon_device_info(int device_id, int* sm_count, int* cc_major, int* cc_minor);

/* ============================ a1-a3: embeddings ================================== */
/* Replaces the 26 Keras `Embedding` gathers (+Cast) of SparseEmbed.call (IL:225-242),
 * the optional field `Add` (IL:233-234) and SeqBaseLayer's sum-pool (BL:45-46).
 *
 *   arena   [R, dim] f32   every field's table back to back (rows of field f start at
 *                          field_row_offset[f]; field_row_offset has n_fields+1 entries
 *                          and lives on the HOST)
 *   ids     [B, F] or [B, F, L]  int32 | int64, compact
 *   out     [B, F, dim] f32; strides of dims 0 and 1 are free (so it may be a window of
 *                          a wider [B, 13+F*dim] concat buffer, CL:49-55), dim 2 compact
 *   L > 1   sums the L rows of a bag in order l = 0..L-1 (BL:46)
 *   flags   KON_EMBED_SUM_FIELDS: out is [B, dim] = sum over f, left to right (IL:233)
 *   oob     optional [1] int32 device counter incremented per out-of-range id (such
 *           lookups return zeros, as TF's GPU gather does; TF-CPU raises)
 */
#define KON_EMBED_SUM_FIELDS 1
int kon_embed_fwd(const DLTensor* arena, const DLTensor* ids, const int64_t* field_row_offset,
                  int32_t n_fields, DLTensor* out, DLTensor* oob, int32_t flags, void* stream);

/* a4: embedding backward (implicit in Model.fit; TF: IndexedSlices -> unique +
 * unsorted_segment_sum).  Sort-then-segment, deterministic.
 *
 *   d_out        [B, F, dim] f32, free strides on dims 0/1 (stride 0 on dim 1 = the
 *                gradient of a field-summed lookup)
 *   ids          as in kon_embed_fwd
 *   unique_rows  [N] int32 out (arena row ids, ascending), N = B*F*L
 *   grads        [N, dim] f32 out; row u is the summed gradient of unique_rows[u]
 *   n_unique     [1] int32 out (device)
 *   workspace    [>= kon_embed_bwd_workspace_bytes(N, dim)] uint8
 */
size_t kon_embed_bwd_workspace_bytes(int64_t n_lookups, int32_t dim);
int kon_embed_bwd(const DLTensor* d_out, const DLTensor* ids, const int64_t* field_row_offset,
                  int32_t n_fields, DLTensor* unique_rows, DLTensor* grads, DLTensor* n_unique,
                  DLTensor* workspace, void* stream);

/* Same as kon_embed_bwd, but skips the routing (per-field counting sort, run-head count) and reuses the sorted
 * (key, position, run id) arrays a previous kon_embed_bwd call left at the front of the SAME
 * `workspace` for the SAME ids and field_row_offset (e.g. the first-order tables after the
 * embedding tables of one step: identical routing, different payload width).  The workspace must be
 * large enough for both calls (max of kon_embed_bwd_workspace_bytes over the two dims). */
int kon_embed_bwd_reuse(const DLTensor* d_out, const DLTensor* ids, const int64_t* field_row_offset,
                        int32_t n_fields, DLTensor* unique_rows, DLTensor* grads, DLTensor* n_unique,
                        DLTensor* workspace, void* stream);

/* Routing only (per-field counting sort, run-head count) into the front of `workspace`; the routing depends on
 * the ids alone, so a trainer can run it on a side stream at the start of the step and call
 * kon_embed_bwd_reuse (or kon_embed_bwd_peer with reuse_sort = 1) with the same workspace in the backward. */
int kon_embed_sort(const DLTensor* ids, const int64_t* field_row_offset, int32_t n_fields,
                   DLTensor* workspace, void* stream);

/* The routing's pass plan for ONE table (no device work; diagnostics and host-side tests): a table of `rows` rows is
 * sorted by a stable LSD counting sort with digits of at most 12 bits, its passes occupying the first slots of the
 * job.  out[6] = {active in `slot`, first pass, last pass, shift, digits, mask}; returns the passes the table needs. */
int kon_embed_route_plan(int64_t rows, int32_t slot, int32_t max_passes, int64_t* out);

/* Two gradients over ONE routing, one pass: d_out [B,F,dim] (dim % 4 == 0) for the embedding arena and
 * d_lin [B,F,1] (any strides; stride_f = 0 for a sum-pooled first-order term) for the dim-1 "linear" arena that
 * FeatureInput(useLinear=True) (DP:65-76) looks up with the SAME ids and per-field row counts.  unique_rows /
 * n_unique are shared; grads [>=N,dim], grads_lin [>=N,1].  reuse_sort != 0: the sorted routing of these ids is
 * already at the front of `workspace` (kon_embed_sort).  Workspace: kon_embed_bwd_workspace_bytes(N, dim). */
int kon_embed_bwd_pair(const DLTensor* d_out, const DLTensor* d_lin, const DLTensor* ids,
                       const int64_t* field_row_offset, int32_t n_fields, DLTensor* unique_rows,
                       DLTensor* grads, DLTensor* grads_lin, DLTensor* n_unique, DLTensor* workspace,
                       int32_t reuse_sort, void* stream);

/* Sparse row-wise SGD on the arena: w[r] -= lr * (g + 2*l2*w[r]) for the n_unique rows
 * (the L2 term is the reference's embeddings_regularizer, IL:217, applied lazily). */
int kon_embed_sgd(DLTensor* arena, const DLTensor* unique_rows, const DLTensor* grads,
                  const DLTensor* n_unique, float lr, float l2, void* stream);
/* Sparse row-wise (lazy) Adam on the arena; m, v are [R, dim] f32 state, step >= 1. */
int kon_embed_adam(DLTensor* arena, DLTensor* m, DLTensor* v, const DLTensor* unique_rows,
                   const DLTensor* grads, const DLTensor* n_unique, float lr, float beta1,
                   float beta2, float eps, float l2, int32_t step, void* stream);
/* Same, with the step counter read from a device int32[1] (the bias correction is computed in the
 * kernel), so that a captured CUDA graph of the training step can be replayed. */
int kon_embed_adam_devstep(DLTensor* arena, DLTensor* m, DLTensor* v, const DLTensor* unique_rows,
                           const DLTensor* grads, const DLTensor* n_unique, float lr, float beta1,
                           float beta2, float eps, float l2, const DLTensor* step, void* stream);

/* ================= 8e: sharded embeddings over NVLink peer memory =================== */
/* The reference is single-process (no distributed code at all); this is the scale-out of
 * SparseEmbed.call (IL:225-242) and of its implicit gradient for one process per GPU.  Tables are
 * sharded table-wise / row-wise over the ranks (ml_function_b200/parallel.py); the pooled-embedding
 * all-to-all of the forward and the dOut all-to-all of the backward are NOT separate collectives:
 * the gather kernel stores each row straight into the concat buffer of the rank that owns the
 * sample, and the segmented reduction loads each gradient row straight from the rank that
 * produced it -- 16-byte accesses over NVLink 5 / NVSwitch through CUDA-IPC mappings.
 *
 * kon_peer_alloc   cudaMalloc + zero + cudaIpcGetMemHandle; `handle64` receives the 64 opaque
 *                  bytes another process passes to kon_peer_open (cudaIpcOpenMemHandle, peer
 *                  access enabled lazily).  kon_peer_close / kon_peer_free undo them.
 * kon_peer_barrier one tiny kernel on `stream`: rank `rank` releases a flag into every peer's flag
 *                  block and acquires every peer's flag in its own; peer_flags[q] = this process's
 *                  mapping of rank q's flag block (>= 128 zero-initialised bytes).  Graph-capturable.
 *                  A peer that does not arrive within timeout_ms (<= 0: 10 s) sets word 17 of the
 *                  caller's own flag block instead of hanging the GPU.
 */
int kon_peer_alloc(int device_id, size_t bytes, void** ptr, void* handle64);
int kon_peer_open(int device_id, const void* handle64, void** ptr);
int kon_peer_close(int device_id, void* ptr);
int kon_peer_free(int device_id, void* ptr);
int kon_peer_barrier(void* const* peer_flags, int32_t n_peers, int32_t rank, int device_id,
                     int64_t timeout_ms, void* stream);

/* The small exchanges of the sharded step as peer STORES (no NCCL call, no staging): up to KON_MAX_PUTS strided
 * 2-D copies in one launch, sources in local memory, destinations anywhere (kon_peer_open mappings).  Used for the
 * all-to-all of the 4-byte ids (columns -> owning rank), of the output gradient rows (columns -> owning rank, so the
 * scatter-add then reads local memory only), and of the first-order partial sums / their gradients.  All byte
 * quantities are multiples of 4; 16-byte units are used when everything is 16-byte aligned.  A kon_peer_barrier on
 * the same stream publishes the data. */
#define KON_MAX_PUTS 64
typedef struct {
  const void* src;      /* local */
  void*       dst;      /* local or peer mapping */
  int64_t     src_pitch, dst_pitch;   /* bytes between rows */
  int64_t     width;    /* bytes per row */
  int64_t     rows;
} KonPut2D;
int kon_peer_put2d(const KonPut2D* puts, int32_t n, int device_id, void* stream);

/* Forward: `ids` [B_global, F_local] are this rank's lookups for the GLOBAL batch (F_local = the
 * fields whose tables this rank owns, `field_row_offset` into its local `arena`).  Sample b belongs
 * to rank q = b / rows_per_peer; its row for local field f is stored at
 *     peer_out[q] + (b - q*rows_per_peer)*out_stride_b + f*out_stride_f        (floats)
 * so peer_out[q] already points at the column of this rank's first field inside rank q's buffer.
 * flags: KON_EMBED_SKIP_INVALID leaves the destination untouched for out-of-range ids (row-wise
 * shards mark the rows of other ranks with -1; their owner writes them). */
#define KON_EMBED_SKIP_INVALID 2
int kon_embed_fwd_peer(const DLTensor* arena, const DLTensor* ids, const int64_t* field_row_offset,
                       int32_t n_fields, void* const* peer_out, int32_t n_peers,
                       int64_t rows_per_peer, int64_t out_stride_b, int64_t out_stride_f,
                       DLTensor* oob, int32_t flags, void* stream);
/* Same, with an explicit destination column per local field: the row of local field f lands at float offset
 * field_col[f] of the destination row (instead of f*out_stride_f), so an owner's fields need not be adjacent in the
 * model's field order (cost-balanced table placement) and the receiver's buffer is still in model order. */
int kon_embed_fwd_peer_cols(const DLTensor* arena, const DLTensor* ids, const int64_t* field_row_offset,
                            int32_t n_fields, void* const* peer_out, int32_t n_peers,
                            int64_t rows_per_peer, int64_t out_stride_b, int64_t out_stride_f,
                            const int32_t* field_col, DLTensor* oob, int32_t flags, void* stream);
/* Backward: kon_embed_bwd with the gradient row of (sample b, local field f) loaded from
 *     peer_d_out[q] + (b - q*rows_per_peer)*stride_b + f*stride_f,   q = b / rows_per_peer.
 * Out-of-range ids (rows owned by another rank) carry no gradient.  reuse_sort != 0: as
 * kon_embed_bwd_reuse (the sorted routing of the same ids is already at the front of `workspace`). */
int kon_embed_bwd_peer(const void* const* peer_d_out, int32_t n_peers, int64_t rows_per_peer,
                       int64_t stride_b, int64_t stride_f, int32_t dim, const DLTensor* ids,
                       const int64_t* field_row_offset, int32_t n_fields, DLTensor* unique_rows,
                       DLTensor* grads, DLTensor* n_unique, DLTensor* workspace, int32_t reuse_sort,
                       void* stream);

/* ============================ a5-a6: FM ========================================== */
/* Replaces InnerLayer's 325 tf.multiply + sequential Add (IL:59-66) and FmLayer's Add
 * of the linear terms (IL:161-170) by one pass:
 *   out[b,:] = sum_{i<j} v[b,i,:]*v[b,j,:] + sum_f lin[b,f]
 *   v [B,F,k] f32 (free strides on dims 0/1), out [B,k] f32,
 *   lin [B,Fl] f32 or NULL: Fl = F for the reference's per-field list; any Fl >= 1 is accepted
 *   (only the sum over lin's fields enters, e.g. an already reduced [B,1]). */
int kon_fm_fwd(const DLTensor* v, const DLTensor* lin, DLTensor* out, void* stream);
/* dv[b,f,:] = g[b,:] * (S[b,:] - v[b,f,:]);  dlin[b,f] = sum_k g[b,k]  (dlin [B,Fl] or NULL) */
int kon_fm_bwd(const DLTensor* v, const DLTensor* g, DLTensor* dv, DLTensor* dlin, void* stream);
/* Same, but dv += ...: `dv` already holds another consumer's gradient of the same buffer (DeepFM: the first Dense
 * layer's input gradient, CL:190, of the concat buffer the FM window is cut from), so the `Add` autograd would run
 * over two [B, 13+F*k] tensors is folded into this pass.  dlin is written, not accumulated. */
int kon_fm_bwd_acc(const DLTensor* v, const DLTensor* g, DLTensor* dv, DLTensor* dlin, void* stream);

/* ============================ a7: DCN cross ====================================== */
/* Replaces CrossLayer.call (IL:275-282): x_{l+1} = x0 * (x_l . w_l) + x_l + b_l.
 *   x0 [B,D] f32, w,b [L,D] f32 (the reference's L `[D,1]` kernels / biases stacked),
 *   out [B,D] f32 (the reference returns it as [B,D,1]),
 *   s [B,L] f32 out: the per-sample scalars x_l . w_l, saved for the backward. */
int kon_cross_fwd(const DLTensor* x0, const DLTensor* w, const DLTensor* b, DLTensor* out,
                  DLTensor* s, void* stream);
size_t kon_cross_bwd_workspace_bytes(int64_t batch, int32_t dim, int32_t layers, int device_id);
int kon_cross_bwd(const DLTensor* x0, const DLTensor* w, const DLTensor* b, const DLTensor* s,
                  const DLTensor* g, DLTensor* dx0, DLTensor* dw, DLTensor* db,
                  DLTensor* workspace, void* stream);
/* Same, but dx0 += ...: `dx0` already holds another consumer's gradient of x0 (DCN: the first Dense layer's input
 * gradient of the concat buffer both branches read, MD:98-101); dw / db are written as usual. */
int kon_cross_bwd_acc(const DLTensor* x0, const DLTensor* w, const DLTensor* b, const DLTensor* s,
                  const DLTensor* g, DLTensor* dx0, DLTensor* dw, DLTensor* db,
                  DLTensor* workspace, void* stream);

/* ============================ a8: xDeepFM CIN ==================================== */
/* Replaces CIN.call (IL:310-327) up to and including the per-layer sum pooling; the
 * Dense(1) logit layer (IL:325) stays with the caller.
 *
 *   x0       [B, m, D] f32 compact           (Concatenate(axis=1) of the field embeddings, MD:131)
 *   w[l]     [H_{l-1}*m, H_l] f32            Keras Conv1D kernel [1,C,N] without its leading 1;
 *                                            channel c = h*m + i (IL:317-318); H_0 = m
 *   bias[l]  [H_l] f32
 *   pooled   [B, n_layers*D] f32 out         concat over layers of sum_o z_l[b,d,o] (IL:322-323)
 *   saved    [>= kon_cin_saved_bytes] uint8  activations kept for the backward (z_l)
 *   precision KON_CIN_FP32: fp32 CUDA-core arithmetic (parity mode, 1e-5)
 *             KON_CIN_BF16: bf16 operands on tcgen05 tensor cores, fp32 accumulate (2e-2)
 */
#define KON_CIN_FP32 0
#define KON_CIN_BF16 1
#define KON_CIN_MAX_LAYERS 8
size_t kon_cin_saved_bytes(int64_t batch, int32_t m, int32_t D, const int32_t* layer_sizes,
                           int32_t n_layers, int32_t precision);
size_t kon_cin_workspace_bytes(int64_t batch, int32_t m, int32_t D, const int32_t* layer_sizes,
                               int32_t n_layers, int32_t precision, int device_id);
int kon_cin_fwd(const DLTensor* x0, const DLTensor* const* w, const DLTensor* const* bias,
                int32_t n_layers, DLTensor* pooled, DLTensor* saved, DLTensor* workspace,
                int32_t precision, void* stream);
/* d_pooled [B, n_layers*D] -> dx0 [B,m,D], dw[l] like w[l], dbias[l] like bias[l]. */
int kon_cin_bwd(const DLTensor* x0, const DLTensor* const* w, const DLTensor* const* bias,
                int32_t n_layers, const DLTensor* d_pooled, const DLTensor* saved,
                DLTensor* dx0, DLTensor* const* dw, DLTensor* const* dbias,
                DLTensor* workspace, int32_t precision, void* stream);

/* ============================ a9-a10: AutoInt attention ========================== */
/* Replaces MultHeadAttentionLayer.call (BL:356-377) + ProductAttentionLayer.call
 * (BL:292-311) + the Add/ReLU the wrapping DnnLayer applies (CL:205-216):
 *   Q = X Wq, K = X Wk, (V = K, BL:360), S = sigmoid(Q K^T [/ sqrt(d)]),
 *   O = S K, Y = ReLU(LN_{eps}(O) * gamma + beta + X Wr)
 *   x [B,F,kin] f32; wq,wk,wr [kin,H,d] f32; gamma,beta [d] f32; y [H,B,F,d] f32.
 * flags select the reference's ctor switches. */
#define KON_ATTN_USE_SCALE 1   /* use_scale (BL:296-297)              */
#define KON_ATTN_USE_LN    2   /* use_ln    (BL:368-369), eps = 1e-3  */
#define KON_ATTN_USE_RES   4   /* use_res   (BL:365-366) + Add (CL:212) */
#define KON_ATTN_RELU      8   /* DnnLayer's activation (CL:216)      */
#define KON_ATTN_BF16      16  /* bf16 operands on tensor cores (warp-level MMA), fp32 accumulate, sigmoid /
                                  LayerNorm in fp32; tolerance 2e-2.  Needs F <= 32, kin in {16,32,48,64},
                                  d == 8.  Without it: fp32 CUDA-core arithmetic (parity mode, 1e-5). */
int kon_attn_fwd(const DLTensor* x, const DLTensor* wq, const DLTensor* wk, const DLTensor* wr,
                 const DLTensor* gamma, const DLTensor* beta, DLTensor* y, float ln_eps,
                 int32_t flags, void* stream);
size_t kon_attn_bwd_workspace_bytes(int64_t batch, int32_t fields, int32_t kin, int32_t heads,
                                    int32_t d, int device_id);
int kon_attn_bwd(const DLTensor* x, const DLTensor* wq, const DLTensor* wk, const DLTensor* wr,
                 const DLTensor* gamma, const DLTensor* beta, const DLTensor* gy, DLTensor* dx,
                 DLTensor* dwq, DLTensor* dwk, DLTensor* dwr, DLTensor* dgamma, DLTensor* dbeta,
                 DLTensor* workspace, float ln_eps, int32_t flags, void* stream);

/* ============================ a12: skinny heads ================================== */
/* Replaces MergeScoreLayer.call (CL:86-100: Flatten + Concatenate + Dense(2); the softmax stays
 * with the caller) and Dense(1) logit layers (CL:190): y = [x1 | x2] W + b without the concat copy.
 *   x1 [B,D1] f32, x2 [B,D2] f32 or NULL (free row strides; with x2: D1 % 4 == 0), D1+D2 <= 1024,
 *   w [D1+D2, N] f32 with N in {1,2} (Keras Dense kernel layout), b [N] or NULL, y [B,N].
 * Backward: dx1 / dx2 (either may be NULL: not needed), dw [D,N], db [N] or NULL; deterministic. */
int kon_head_fwd(const DLTensor* x1, const DLTensor* x2, const DLTensor* w, const DLTensor* b,
                 DLTensor* y, void* stream);
size_t kon_head_bwd_workspace_bytes(int64_t batch, int32_t dim, int32_t units, int device_id);
int kon_head_bwd(const DLTensor* x1, const DLTensor* x2, const DLTensor* w, const DLTensor* gy,
                 DLTensor* dx1, DLTensor* dx2, DLTensor* dw, DLTensor* db, DLTensor* workspace,
                 void* stream);

/* ===================== 8f: callers / siblings wired from the same layer classes ================ */
/* SeqBaseLayer.call on a materialised sequence embedding (BL:45-46: reduce_sum(input_, axis=1)) and the
 * reduce_sum over the pair axis in AFM's AttentionBaseLayer (IL:364):
 *   x [B,L,k] f32 (free strides on dims 0/1) -> out [B,k] = sum_l x[b,l,:], l = 0..L-1 in order.
 * (The fused gather + pool of a sequence feature is kon_embed_fwd with [B,F,L] ids.) */
int kon_pool_sum_fwd(const DLTensor* x, DLTensor* out, void* stream);

/* InnerLayer.call with use_inner=True, use_add=False (IL:61): the un-summed list of Hadamard products
 * [v_i * v_j for i<j] in itertools.combinations order, packed as out [B, F(F-1)/2, k]
 * (AFM, MD:143; IPNN, IL:68-80).  v [B,F,k] f32 (free strides on dims 0/1), 2 <= F <= 64.
 * Backward: dv[b,f,:] = sum_{j != f} g[b,pair(f,j),:] * v[b,j,:]. */
int kon_pairs_fwd(const DLTensor* v, DLTensor* out, void* stream);
int kon_pairs_bwd(const DLTensor* v, const DLTensor* g, DLTensor* dv, void* stream);

/* ProductAttentionLayer.call (BL:292-311) on explicit [q, k, v] (each [..., F, d] f32 compact, identical
 * shapes; leading dims are batch): out = sigmoid(mask(q k^T [/ sqrt(d)])) v.
 *   mask_mode 0: no mask;  1: score = score @ mask (BL:300-302);  2: score += mask * (-100000) (BL:303-306);
 *   mask [F,F] f32 (the reference casts its boolean mask to float the same way).
 * Backward recomputes the scores: dq, dk, dv like q. */
int kon_pattn_fwd(const DLTensor* q, const DLTensor* k, const DLTensor* v, const DLTensor* mask,
                  DLTensor* out, int32_t use_scale, int32_t mask_mode, void* stream);
int kon_pattn_bwd(const DLTensor* q, const DLTensor* k, const DLTensor* v, const DLTensor* mask,
                  const DLTensor* g_out, DLTensor* dq, DLTensor* dk, DLTensor* dv, int32_t use_scale,
                  int32_t mask_mode, void* stream);

/* compile(loss=binary_crossentropy) on probabilities (example/ctr_example/un_seq.py:61), forward and backward in
 * one kernel each instead of ~30 elementwise launches:
 *   p' = clip(p, eps, 1-eps);  loss = mean over all elements of -(y log(p'+eps) + (1-y) log(1-p'+eps))
 *   p, y compact f32 of one size ([B,2] softmax heads with one-hot labels, [B,1,1] sigmoid heads); loss f32[1];
 *   workspace >= kon_bce_workspace_bytes() uint8 (per-CTA partials, summed in fixed order: deterministic).
 * Backward: dp = g_loss * dloss/dp (zero where the clip is active, as tf.clip_by_value). */
size_t kon_bce_workspace_bytes(void);
int kon_bce_fwd(const DLTensor* p, const DLTensor* y, DLTensor* loss, DLTensor* workspace, float eps, void* stream);
int kon_bce_bwd(const DLTensor* p, const DLTensor* y, const DLTensor* g_loss, DLTensor* dp, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KON_B200_H_ */
