/*
 * srb200.h — C ABI of libsrb200.so, the B200 (sm_100a) kernel library behind the SRModel
 * plugin API of george-gca/sr-pytorch-lightning.
 *
 * The reference has no FFI of its own: its hot path is Python calling torch.nn modules
 * (SURVEY.md §8b).  The entry points below are what a binding for that path replaces; each
 * one cites the reference call site(s) it stands in for.  Rules of the ABI:
 *   - plain pointers and sizes only (device pointers come from tensor.data_ptr(); the stream is
 *     a cudaStream_t passed as void*); no torch types; POD structs;
 *   - every function returns 0 on success, non-zero on error; srb_last_error() gives the text
 *     (thread-local); nothing throws or exits across the boundary;
 *   - every kernel is launched on the given stream, never synchronises the host and never
 *     allocates device memory, so every call is CUDA-graph capturable;
 *   - activations are NHWC ("pixel-major") with an explicit channel stride `cs` (elements per
 *     pixel) and channel offset `co`, which is how RDN's torch.cat (models/rdn.py:21,108) is
 *     eliminated: a dense block is one [N,H,W,576] buffer and every conv reads a channel prefix
 *     and writes a channel slice of it.
 */
#ifndef SRB200_H
#define SRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRB_ABI_VERSION 1

typedef struct srb_ctx srb_ctx;

/* element type of activation tensors */
enum { SRB_F32 = 0, SRB_BF16 = 1 };

/* conv epilogue flags:  v = acc (+bias); RELU; v *= scale; MASK: v = mask>0 ? v : 0;
 *                       RESIDUAL: v += residual; store (optionally pixel-shuffled); COLSUM */
enum {
  SRB_RELU     = 1,   /* nn.ReLU(True) after the conv (common.py:46,89; rcan.py:36,46; rdn.py:16) */
  SRB_RESIDUAL = 2,   /* `res += x` (common.py:107; rcan.py:54,73,122; edsr.py:47; rdn.py:40,109) */
  SRB_MASK     = 4,   /* ReLU backward: zero where the saved post-ReLU activation is <= 0          */
  SRB_COLSUM   = 8,   /* per-channel sums of the stored values: CALayer's AdaptiveAvgPool2d
                         numerator (rcan.py:14,25) in forward, bias gradients in backward          */
  SRB_OUT2     = 16   /* also store the result into a second tensor (RDN: LFF output feeds both the
                         next RDB and the GFF concat buffer, rdn.py:104-108)                        */
};

/* weight packings produced by srb_pack_weight */
enum {
  SRB_PACK_SIMT = 0,  /* fp32 [kh][kw][Cin][Cout]                        (CUDA-core kernels)       */
  SRB_PACK_UMMA = 1   /* bf16 [Cin/64][kw][kh][Cout][64], K-major 128-B rows (tcgen05 kernels)     */
};
enum { SRB_PACK_FWD = 0, SRB_PACK_DGRAD = 1 };

/* kernel family selection */
enum { SRB_BACKEND_AUTO = 0, SRB_BACKEND_SIMT = 1, SRB_BACKEND_UMMA = 2 };

typedef struct srb_conv_desc {
  int32_t N, H, W;          /* conv input == conv output spatial size (stride 1, padding k/2)    */
  int32_t Cin, Cout, ksize; /* ksize odd; tcgen05 path: 1 or 3                                     */
  int32_t dtype;            /* SRB_F32 / SRB_BF16 for x, y, residual, mask, y2                      */
  int32_t flags;
  float   scale;            /* res_scale (common.py:106); 1.0 when unused                          */
  int32_t shuffle;          /* 0, or r (2|3): fold nn.PixelShuffle(r) (common.py:133; rdn.py:87-92)
                               into the store: y is [N, H*r, W*r, Cout/(r*r)]                      */
  int32_t colsum_groups;    /* with SRB_COLSUM: 1 -> colsum[Cout]; N -> colsum[N][Cout]            */
  int32_t backend;
  int32_t x_cs, x_co;       /* channel stride / offset (elements) of x                             */
  int32_t y_cs, y_co;       /* ... of y (in the shuffled tensor when shuffle != 0)                  */
  int32_t r_cs, r_co;       /* ... of residual (same pixel grid as y)                               */
  int32_t m_cs, m_co;       /* ... of mask     (same pixel grid as y)                               */
  int32_t y2_cs, y2_co;     /* ... of y2                                                            */
} srb_conv_desc;

typedef struct srb_wgrad_desc {
  int32_t N, H, W;
  int32_t Cin, Cout, ksize;
  int32_t dtype;            /* of x and gy */
  int32_t accumulate;       /* 0: overwrite dw/dbias, 1: += (gradient accumulation)                */
  int32_t shuffle;          /* r if gy is given in the un-shuffled (i,j,c') channel order          */
  int32_t backend;
  int32_t x_cs, x_co;
  int32_t g_cs, g_co;
  float   alpha;            /* dw = alpha * (x (*) gy): res_scale of common.py:106 on the weight grad */
} srb_wgrad_desc;

/* ---- lifetime ---------------------------------------------------------------------------- */
int  srb_abi_version(void);
const char* srb_last_error(void);
int  srb_create(int device, srb_ctx** out);
int  srb_destroy(srb_ctx* ctx);
int  srb_num_sms(const srb_ctx* ctx);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long srb_launch_count(void);

/* ---- weights ------------------------------------------------------------------------------
 * nn.Conv2d.weight is fp32 OIHW (state_dict contract).  Packed copies are caches.
 * mode DGRAD packs the 180-degree-rotated, channel-swapped filter so that the input gradient of
 * a conv is computed by the forward kernel (autograd's convolution_backward for common.py:7-30).
 * shuffle = r permutes output channels from (c', i, j) to (i, j, c') order (see srb_conv_desc). */
size_t srb_packed_weight_bytes(int Cout, int Cin, int ksize, int packing, int mode);
int  srb_pack_weight(srb_ctx*, const float* w_oihw, int Cout, int Cin, int ksize, int packing, int mode,
                     int shuffle, void* out, void* stream);
/* bias permuted the same way as the packed output channels (fp32 [Cout]) */
int  srb_pack_bias(srb_ctx*, const float* bias, int Cout, int shuffle, float* out, void* stream);
/* Re-pack MANY weights in one launch (after an optimizer step): `table` is a DEVICE array of n
 * items, built once because parameter and packed-buffer addresses are stable.  ksize == 0 marks a
 * bias item (srb_pack_bias semantics).  max_elems = the Cout*Cin*k*k the grid rows are sized for (any value >= 1 is
 * correct: items are processed with grid-stride loops; the median filter size of the table is the efficient choice). */
typedef struct srb_pack_item {
  const float* src;
  void* dst;
  int32_t Cout, Cin, ksize, packing, mode, shuffle;
} srb_pack_item;
int  srb_pack_table(srb_ctx*, const srb_pack_item* table_dev, int n, int64_t max_elems, void* stream);

/* ---- convolution --------------------------------------------------------------------------
 * Replaces nn.Conv2d.forward at every call site of models/common.py:7-30 (DefaultConv2d),
 * rcan.py:41-42,68,101, rdn.py:15,37,59,71-72,87-93, edsr.py:21-33, with the elementwise ops that
 * follow it in the reference fused into the epilogue.  Also computes autograd's input gradient
 * when given DGRAD-packed weights (bias = NULL). */
int  srb_conv(srb_ctx*, const srb_conv_desc*, const void* x, const void* w_packed, const float* bias,
              const void* residual, const void* mask, void* y, void* y2, float* colsum, void* stream);

/* autograd's weight / bias gradient of the same conv: dw is fp32 OIHW [Cout][Cin][k][k],
 * dbias fp32 [Cout] (may be NULL). */
int  srb_conv_wgrad(srb_ctx*, const srb_wgrad_desc*, const void* x, const void* gy,
                    float* dw_oihw, float* dbias, void* stream);
/* Many weight gradients in one call (the backward pass defers them: they are off the critical
 * path of the input-gradient chain).  Eligible layers share batched tcgen05 launches. */
typedef struct srb_wgrad_item {
  srb_wgrad_desc d;
  const void* x;
  const void* gy;
  float* dw;
  float* dbias;             /* may be NULL */
} srb_wgrad_item;
/* an item with dw == NULL and dbias != NULL is a bias gradient only (column sums of gy; x is ignored) */
int  srb_conv_wgrad_batched(srb_ctx*, const srb_wgrad_item* items, int n, void* stream);
/* max_ctas > 0: the batched weight-gradient launches of later srb_conv_wgrad_batched calls use at most that many CTAs
 * (one per SM), so that they fit beside a kernel that leaves SMs free — the per-sample cluster chain (96 of 148 SMs for
 * 16 x 48x48) on another stream; 0 restores "all SMs".  Replaces nothing in the reference: torch runs the weight
 * gradients of a layer right behind its input gradient on the same stream (autograd engine). */
int  srb_set_wgrad_sm_budget(srb_ctx*, int max_ctas);
/* a one-thread kernel that occupies the stream for ns nanoseconds (<= 1 ms): lets a kernel on ANOTHER stream that becomes
 * runnable at the same moment reach the SMs first (the cluster chain before the weight gradients that run beside it) */
int  srb_delay(srb_ctx*, int64_t ns, void* stream);
/* which kernel family (and hence which weight packing) backend AUTO resolves to */
int  srb_conv_uses_umma(const srb_conv_desc*);
int  srb_wgrad_uses_umma(const srb_wgrad_desc*);
/* Diagnostics (no device needed): how srb_conv_wgrad_batched groups n tcgen05 weight-gradient blocks
 * (64 ci x 64 co x taps of one layer; ntiles[i] 128-pixel tiles, k1[i] != 0 for a 1x1 layer) into launches
 * on a device with num_sms SMs.  Outputs in plan order (sorted by cost): tiles_out[i] (negative: 1x1 block),
 * launch_out[i], ctas_out[i] = CTAs that share block i (each takes every ctas_out[i]-th tile). */
int  srb_wgrad_plan(int num_sms, int n, const int* ntiles, const int* k1, int* tiles_out, int* launch_out, int* ctas_out);

/* ---- layer chains ---------------------------------------------------------------------------
 * MANY dependent 64-channel layers in ONE persistent launch.  A single 3x3 64->64 conv on a
 * 16 x 48x48 batch is ~1.8 us of tensor work but ~6.5 us as a dependent kernel (launch gap,
 * prologue, filter reload, L2 round trip).  A chain runs a whole ResidualGroup (models/rcan.py:59-74:
 * 20 x RCAB + conv + skip) or a stack of ResBlocks (common.py:74-109), forward or backward, as one
 * kernel: every CTA owns fixed output tiles, filters stream through shared memory layer by layer,
 * and per-(op, sample) counters in device memory order the layers; tiles of DIFFERENT samples
 * interleave on an SM so one sample's store -> flag -> load latency hides behind another's MMAs.
 *
 * Tensors are slots of up to four "spaces", each a contiguous [slots][N][H][W][64] bf16 buffer;
 * a buffer reference is (space << 14) | slot, SRB_CHAIN_NONE when unused.  Every slot is written
 * by at most one op of a chain.  Op i (i > 0) may read what ops < i wrote; op 0 reads only data
 * produced before the launch.
 *
 * SRB_CHAIN_CONV: y = epilogue(conv3x3(x)) with the srb_conv flag semantics (RELU, scale, MASK or
 *   RESIDUAL tile `e`, COLSUM).  With SRB_CHAIN_CA the op is an RCAB's second conv fused with its
 *   CALayer (rcan.py:10-29,54): t = conv(x)+bias is stored to y, its per-sample channel sums go to
 *   colsum [N][64]; once the sample's sums are complete, out = t*gate + e is stored to y2 and the
 *   pooled mean / gate are written to ca_s / ca_y [N][64] for backward.
 * SRB_CHAIN_CA_BWD: CALayer + skip backward for one RCAB: x = t (saved), e = g (dL/dout);
 *   y = dt = g*gate + ds/HW; parameter gradients are ACCUMULATED into ca_dw1..ca_db2, column sums
 *   of dt into colsum [64] (the bias gradient of the conv that produced t); ca_scratch [N][64]
 *   must be zero on entry.
 * counters: n_ops * 2 * N int32, zero on entry. */
enum { SRB_CHAIN_CONV = 0, SRB_CHAIN_CA_BWD = 1 };
enum { SRB_CHAIN_CA = 32,              /* extra flag bits for SRB_CHAIN_CONV ops */
       SRB_CHAIN_CA_BWD_FUSED = 64,    /* after y is stored, run CA_BWD on it: g = y, t = tile e2, dt -> y2,
                                          column sums of dt -> colsum2 (saves a whole dependent op per RCAB) */
       SRB_CHAIN_Y_SCRATCH = 128 };    /* hint: nothing outside this chain reads slot y; a kernel that hands y to its
                                          consumers on chip (conv_cluster.cu: TMEM-parked residual) may skip the store */
#define SRB_CHAIN_NONE 0xFFFFu
#define SRB_CHAIN_MAX_OPS 64

typedef struct srb_chain_op {
  int32_t  kind;
  uint32_t flags;
  uint16_t x, y, e, y2;     /* buffer references */
  uint16_t e2, reserved0;   /* second operand tile (SRB_CHAIN_CA_BWD_FUSED: the saved t) */
  int32_t  w_layer;         /* index into the packed filter bank (CONV) */
  float    scale;
  int32_t  colsum_groups;   /* 1 -> colsum[64]; N -> colsum[N][64] (always N with SRB_CHAIN_CA) */
  int32_t  ca_cr;           /* hidden width of the CALayer (channel / reduction) */
  float    colsum_scale;    /* factor on the COLSUM contribution (res_scale on a bias gradient); 0 means 1 */
  const float* bias;        /* [64] or NULL */
  float*   colsum;
  float*   colsum2;         /* SRB_CHAIN_CA_BWD_FUSED: [64] column sums of dt, or NULL */
  const float *ca_w1, *ca_b1, *ca_w2, *ca_b2;   /* conv_du.0 [Cr][64], [Cr]; conv_du.2 [64][Cr], [64] */
  float   *ca_s, *ca_y;                         /* [N][64]: written by SRB_CHAIN_CA, read by CA_BWD */
  float   *ca_dw1, *ca_db1, *ca_dw2, *ca_db2;   /* CA_BWD: accumulated */
  float   *ca_scratch;                          /* CA_BWD: [N][64] zero-filled */
} srb_chain_op;

typedef struct srb_chain_desc {
  int32_t N, H, W;                    /* 64 channels, bf16 */
  int32_t n_ops;
  const srb_chain_op* ops;            /* host array */
  void*   space_base[4];
  int32_t space_slots[4];             /* 0 = space unused */
  const void* weights;                /* SRB_PACK_UMMA filters, [n_layers][9][64][64] bf16 */
  int32_t n_layers;
  int32_t* counters;
  int64_t* trace;                     /* diagnostics or NULL: [grid][2][n_ops][8] event times (ns) + [grid][4] (ns, clock) at start/end, see conv_chain.cu */
  int32_t* tile_flags;                /* NULL: layers are ordered by the per-sample counters; else n_ops * N * tiles int32, zero on
                                         entry (tiles = ceil(H/16) * ceil(W/8)): a tile waits only for the 3x3 neighbourhood of
                                         tiles of the previous op */
  int32_t kernel_hint;                /* 0: the cluster kernel where it is eligible (see srb_conv_chain_uses_cluster); 1: always the
                                         L2-flag kernel (all SMs, faster for forward chains that run alone); 2 = 0 */
} srb_chain_desc;
int  srb_conv_chain(srb_ctx*, const srb_chain_desc*, void* stream);
/* 1 if srb_conv_chain runs this chain with the per-sample thread-block-cluster kernel (conv_cluster.cu: H % 16 == 0,
 * W in {24, 48}, at most 8 CTAs of 16 x 24 pixels per sample, every op a conv that consumes the previous op's result;
 * activations stay in shared memory between layers, halos travel through distributed shared memory), 0 if it takes the
 * L2-flag kernel (conv_chain.cu).  SRB200_CHAIN_CLUSTER=0 disables the cluster kernel. */
int  srb_conv_chain_uses_cluster(const srb_chain_desc*);
/* number of CTAs srb_conv_chain launches for this shape (size of the trace buffer's first dim) */
int  srb_conv_chain_grid(const srb_ctx*, int N, int H, int W);

/* ---- RCAN channel attention (models/rcan.py:10-29 CALayer + rcan.py:54 `res += x`) ---------
 * out = t * sigmoid(W2 relu(W1 mean_hw(t) + b1) + b2) + skip.
 * pooled_sum: [N][C] fp32 sums of t over H*W: an input (from srb_conv's COLSUM) when
 * compute_pool == 0, filled here first when compute_pool != 0.
 * s_out [N][C] (mean) and y_out [N][C] (gate) are saved for backward. skip may be NULL. */
int  srb_ca_fwd(srb_ctx*, int N, int H, int W, int C, int Cr, int dtype,
                const void* t, const void* skip, float* pooled_sum, int compute_pool,
                const float* w1, const float* b1, const float* w2, const float* b2,
                void* out, float* s_out, float* y_out, void* stream);
/* backward: g = dL/dout.  dt = g*y + broadcast(ds)/HW; parameter grads are accumulated (+=)
 * when accumulate != 0, else overwritten.  colsum_dt (nullable): per-channel sums of dt over
 * N,H,W = bias gradient of the conv that produced t.  scratch: [N][C] fp32, zero-filled here
 * unless scratch_is_zero (the caller hands out slices of one arena it cleared once per step). */
int  srb_ca_bwd(srb_ctx*, int N, int H, int W, int C, int Cr, int dtype,
                const void* g, const void* t, const float* s, const float* y,
                const float* w1, const float* b1, const float* w2, const float* b2,
                void* dt, float* dw1, float* db1, float* dw2, float* db2,
                float* colsum_dt, float* scratch, int scratch_is_zero, int accumulate, void* stream);

/* ---- boundary layout conversion (model input / output only) --------------------------------
 * NCHW fp32 (what SRModel.forward receives, srmodel.py:163) <-> NHWC dtype, with MeanShift
 * (common.py:58-71, W = I) folded in as a per-channel add (chan_add may be NULL). */
int  srb_nchw_to_nhwc(srb_ctx*, const float* x, int N, int C, int H, int W, const float* chan_add,
                      int dtype, void* y, int y_cs, int y_co, void* stream);
int  srb_nhwc_to_nchw(srb_ctx*, const void* x, int x_cs, int x_co, int dtype, int N, int C, int H, int W,
                      const float* chan_add, float* y, void* stream);

/* ---- small NHWC helpers --------------------------------------------------------------------*/
/* dst[.., d_co:d_co+C] = src[.., s_co:s_co+C]  (npix pixels) */
int  srb_copy_channels(srb_ctx*, const void* src, int s_cs, int s_co, void* dst, int d_cs, int d_co,
                       int C, int64_t npix, int dtype, void* stream);
/* out = a + b over channel slices (gradient accumulation at skip connections) */
int  srb_add_channels(srb_ctx*, const void* a, int a_cs, int a_co, const void* b, int b_cs, int b_co,
                      void* out, int o_cs, int o_co, int C, int64_t npix, int dtype, void* stream);
/* ReLU backward over channel slices: out = act > 0 ? g : 0 (autograd of nn.ReLU, common.py:46) */
int  srb_relu_bwd(srb_ctx*, const void* g, int g_cs, int g_co, const void* act, int a_cs, int a_co,
                  void* out, int o_cs, int o_co, int C, int64_t npix, int dtype, void* stream);
/* adjoint of nn.PixelShuffle(r): g [N,H*r,W*r,C'] -> out [N,H,W,r*r*C'] in (i,j,c') channel order */
int  srb_pixel_unshuffle(srb_ctx*, const void* g, int g_cs, int g_co, void* out, int o_cs, int o_co,
                         int N, int H, int W, int Cp, int r, int dtype, void* stream);
/* per-channel sums over npix pixels: out[C] (+= if accumulate) */
int  srb_colsum(srb_ctx*, const void* x, int x_cs, int x_co, int C, int64_t npix, int dtype,
                float* out, int accumulate, void* stream);

/* ---- loss (models/srmodel.py:37,549 nn.L1Loss; its autograd seed gradient) ------------------
 * loss[0] = mean|sr-hr| ; grad = sign(sr-hr)/n  (both NCHW fp32, n elements). grad may be NULL. */
int  srb_l1_loss(srb_ctx*, const float* sr, const float* hr, int64_t n, float* loss, float* grad,
                 void* stream);

/* ---- optimizer (models/srmodel.py:57,145-154 optim.Adam over a flat fp32 buffer) ----------- */
/* step: 1-based step count; if step_dev != NULL the count is read from device memory instead
 * (so a captured CUDA graph can be replayed; bump it with srb_inc_counter in the same graph).
 * grad_scale multiplies the gradient first (1/world_size for a summed all-reduce). */
int  srb_adam_step(srb_ctx*, float* param, const float* grad, float* m, float* v, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                   const int32_t* step_dev, float grad_scale, void* stream);
int  srb_inc_counter(srb_ctx*, int32_t* counter, void* stream);

/* ---- spatial tiling of large-image inference (BASELINE.json configs[4]; the reference runs the whole frame through
 * SRModel.predict_step on one device, models/srmodel.py:375-380) -----------------------------------------------------
 * Each rank owns a row strip; a layer's output is [1, t + rows + t, W, C] with t halo rows on either side.  One launch per
 * layer copies this rank's first / last t owned rows into the neighbours' halo rows through NVLink peer memory, stores the
 * frame number into the neighbours' flag slot (release, system scope) and waits for the neighbours' flags (acquire), so
 * the next conv on the stream reads complete halos; capturable in a CUDA graph.  Before pushing, the ranks shake hands
 * ("my buffer of this layer is complete, you may write its halo rows": the conv before this call wrote them too).
 * Peer pointers come from srb_ipc_open. */
typedef struct srb_halo_desc {
  const void* src_top;      /* first t owned rows of this rank's buffer            */
  const void* src_bot;      /* last t owned rows                                   */
  void*       dst_up;       /* upper neighbour's bottom halo rows (peer) or NULL   */
  void*       dst_dn;       /* lower neighbour's top halo rows (peer) or NULL      */
  int64_t     slab_bytes;   /* t * W * Cs * elem_size, multiple of 16              */
  int64_t*    flag_up;      /* upper neighbour's "from below" slot pair {ready, data} of this layer (peer) or NULL */
  int64_t*    flag_dn;      /* lower neighbour's "from above" pair (peer) or NULL  */
  const int64_t* wait_up;   /* own "from above" pair (the upper neighbour writes it) or NULL */
  const int64_t* wait_dn;   /* own "from below" pair or NULL                       */
  const int64_t* frame;     /* device counter: current frame number, >= 1 (bump it with srb_inc_counter64 once per frame) */
  uint32_t*   done;         /* zero-initialised device word (CTA completion counter, self-resetting) */
} srb_halo_desc;
int  srb_halo_exchange(srb_ctx*, const srb_halo_desc*, void* stream);
int  srb_inc_counter64(srb_ctx*, int64_t* counter, void* stream);
/* device memory other processes on this node can map (cudaMalloc + cudaIpcGetMemHandle; zero-filled) */
int  srb_ipc_alloc(srb_ctx*, size_t bytes, void** ptr, unsigned char handle[64]);
int  srb_ipc_open(srb_ctx*, const unsigned char handle[64], void** ptr);
int  srb_ipc_close(srb_ctx*, void* ptr);
int  srb_ipc_free(srb_ctx*, void* ptr);


/* ---- BatchNorm2d + PReLU for the SRResNet family (reference models/srresnet.py:9-36 over
 * common.py:33-55,74-109 with norm=nn.BatchNorm2d(n_feats), act=nn.PReLU()).  NHWC, 16-byte channel vectors: C, strides and
 * offsets multiples of 8 (bf16) / 4 (fp32).
 * srb_bn_stats: batch statistics over npix pixels (two passes: mean, then squared deviations): mean[C], rstd[C] =
 *   1/sqrt(biased var + eps); running_mean / running_var (nullable) updated as torch does (momentum, unbiased variance).
 *   ws: 2*C floats of scratch.
 * srb_bn_act_fwd: y = PReLU_a((x - mean) * rstd * gamma + beta) + res; the four BN vectors are given together or all NULL
 *   (no normalisation), prelu_a (one fp32 slope) and res are optional.
 * srb_bn_act_bwd: g = dL/dy -> dx (the residual's gradient is g itself); dgamma / dbeta [C] and da [1] are written, or
 *   added to if accumulate != 0.  ws: 3*C floats of scratch. */
int  srb_bn_stats(srb_ctx*, const void* x, int x_cs, int x_co, int C, int64_t npix, int dtype, float eps, float momentum,
                  float* ws, float* mean, float* rstd, float* running_mean, float* running_var, void* stream);
int  srb_bn_act_fwd(srb_ctx*, const void* x, int x_cs, int x_co, int C, int64_t npix, int dtype, const float* mean,
                    const float* rstd, const float* gamma, const float* beta, const float* prelu_a, const void* res, int r_cs,
                    int r_co, void* y, int y_cs, int y_co, void* stream);
int  srb_bn_act_bwd(srb_ctx*, const void* g, int g_cs, int g_co, const void* x, int x_cs, int x_co, int C, int64_t npix,
                    int dtype, const float* mean, const float* rstd, const float* gamma, const float* beta,
                    const float* prelu_a, float* ws, void* dx, int dx_cs, int dx_co, float* dgamma, float* dbeta, float* da,
                    int accumulate, void* stream);

/* ---- GPU data path for training batches (replaces the per-sample PIL work of the reference's
 * srdata.py:57-169 `_get_item` / `_get_patch`: aligned random LR / HR crop, rotation by a multiple of 90 degrees, flips,
 * TF.to_tensor) for uint8 HWC RGB images resident in device memory.  The host draws the random choices (as the reference
 * does, from Python's `random`) and uploads n items; one launch writes lr_out [n][3][p][p] and hr_out [n][3][p*s][p*s]
 * (fp32, values k/255).  Order of operations and edge behaviour as the reference: crop (zero where the box leaves the
 * image), rotate counter-clockwise, hflip, vflip.  Either output may be NULL. */
typedef struct srb_patch_item {
  const uint8_t* lr_img;    /* [lr_h][lr_w][3] */
  const uint8_t* hr_img;    /* [hr_h][hr_w][3] */
  int32_t lr_h, lr_w, hr_h, hr_w;
  int32_t lr_top, lr_left;  /* crop origin in the LR image; the HR origin is scale times it */
  int32_t angle;            /* 0, 90, 180, 270 */
  int32_t hflip, vflip;
  int32_t reserved;
} srb_patch_item;
int  srb_patch_batch(srb_ctx*, const srb_patch_item* items_dev, int n, int lr_patch, int scale, float* lr_out, float* hr_out,
                     void* stream);

/* per-CTA event clocks of the 3x3 64-channel conv kernel (16 int64 per CTA; NULL = off) */
int  srb_debug_set_trace(srb_ctx*, long long* dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* SRB200_H */
