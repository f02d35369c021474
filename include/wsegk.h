/* wsegk — C ABI of the B200 (sm_100a) kernels behind wseg's Step-1 contrast hot path.
 *
 * The reference (obeychoi0120/wseg) is pure PyTorch and has NO FFI for this path; its only
 * plugin point is `--network <module>` (contrast_train.py:40,60; contrast_infer.py:21,32) and the loss is
 * an inline script body (contrast_train.py:129-395).  Each entry point below therefore cites the chain of
 * ATen calls in the reference that it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory unless named `host_*`;
 *  - tensors are contiguous fp32 NCHW exactly as PyTorch lays them out; indices/labels are int32;
 *  - the caller owns every buffer (inputs, outputs, workspace); nothing is allocated, freed or retained;
 *  - all work is enqueued on `stream` (a CUstream / cudaStream_t); no host synchronisation; graph-capturable;
 *  - return 0 on success or a negative wsegk_status; wsegk_last_error() gives a thread-local message;
 *  - there is no CPU path: on a device that is not sm_100 every call returns WSEGK_ERR_DEVICE.
 */
#ifndef WSEGK_H_
#define WSEGK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* wsegk_stream_t;

enum wsegk_status {
    WSEGK_OK = 0,
    WSEGK_ERR_BAD_SHAPE = -1,
    WSEGK_ERR_BAD_ALIGN = -2,
    WSEGK_ERR_WORKSPACE_TOO_SMALL = -3,
    WSEGK_ERR_CUDA = -4,
    WSEGK_ERR_DEVICE = -5,
    WSEGK_ERR_NULL = -6
};

/* operations, for wsegk_workspace_bytes() */
enum wsegk_op {
    WSEGK_OP_CAM_NORM = 1,
    WSEGK_OP_PCM = 2,        /* one workspace shared by pcm_fwd and pcm_bwd (bwd reads what fwd left) */
    WSEGK_OP_FULLRES_NORM = 3,
    WSEGK_OP_SEED_LABELS = 4,
    WSEGK_OP_PROTO = 5,
    WSEGK_OP_CONTRAST_LOSS = 6,
    WSEGK_OP_SELECT = 7,     /* select_kth and minpool_fwd: N = rows */
    WSEGK_OP_ER_ECR = 8,
    WSEGK_OP_LAZY_BWD = 9    /* fullres_norm_up_bwd and minpool_up_bwd: h, w = the FULL resolution H, W.  Sized for all N*C
                              * planes; only the planes of labelled classes are ever touched. */
};

int wsegk_version(void);
const char* wsegk_last_error(void);

/* Bytes of workspace `op` needs.  Unused dims may be 0.
 *   N batch, C classes (21), K PCM channels (192), D projection dim, h,w grid, P = pixels on the loss grid. */
size_t wsegk_workspace_bytes(int op, int N, int C, int K, int D, int h, int w);

/* K1a — replaces network/resnet38_contrast.py:41-48 (relu, per-(n,c) max-norm, background = 1-max fg,
 * keep only the arg-max foreground class).  cam [N,C,h,w] -> cam_norm [N,C,h,w].  Needs C >= 2. */
int wsegk_cam_norm_fwd(const float* cam, float* cam_norm, int N, int C, int h, int w,
                       void* workspace, size_t workspace_bytes, wsegk_stream_t stream);

/* K2 forward — replaces network/resnet38_contrast.py:70-73 (PCM after the f9 conv):
 *   fh = f/(|f|_2+1e-5); aff = relu(fh^T fh); aff /= colsum(aff)+1e-5; cam_rv = cam_norm @ aff.
 * f [N,K,h,w], cam_norm [N,C,h,w] -> cam_rv [N,C,h,w].  The HWxHW affinity lives only in TMEM.
 * FP16 tensor-core operands, FP32 accumulation; in precise mode (see wsegk_pcm_set_mode) the entries of S too close to
 * zero for their sign to survive the FP16 rounding are re-evaluated in fp32 before the sign mask is handed to pcm_bwd.
 * K % 32 == 0, K <= 192, C <= 23.  The workspace is left holding what pcm_bwd needs (normalised features,
 * norms, column sums and the sign bits of S). */
int wsegk_pcm_fwd(const float* f, const float* cam_norm, float* cam_rv, int N, int C, int K, int h, int w,
                  void* workspace, size_t workspace_bytes, wsegk_stream_t stream);

/* K2 forward when NO backward will follow (contrast_infer.py, torch.no_grad): no sign mask is produced (only pcm_bwd
 * consumes it); cam_rv is identical.  The workspace it leaves must not be handed to pcm_bwd. */
int wsegk_pcm_fwd_infer(const float* f, const float* cam_norm, float* cam_rv, int N, int C, int K, int h, int w,
                        void* workspace, size_t workspace_bytes, wsegk_stream_t stream);

/* Several forward-only K2 problems of different sizes in ONE launch of the tcgen05 kernel (multi-scale +
 * flip inference, contrast_infer.py:58-71 — the 4 scales x {image, flip} of one image are 4 problems of batch 2).  Item i:
 * f[i] [N[i],K,h[i],w[i]], cam_norm[i] [N[i],C,h[i],w[i]] -> cam_rv[i] (same shape as cam_norm[i]); the arrays f, cam_norm,
 * cam_rv, N, h, w, workspaces, workspace_bytes are HOST arrays of n_items (<= 8) entries.  Every item needs its own
 * workspace of the size wsegk_pcm_multi_workspace_bytes reports for it (it depends on how the grid is split between the
 * items).  Results equal wsegk_pcm_fwd_infer's to fp32 rounding (the stage width is 128 pixels for every item). */
int wsegk_pcm_multi_workspace_bytes(int n_items, const int* N, const int* h, const int* w, int C, int K, size_t* bytes_out);
int wsegk_pcm_fwd_infer_multi(int n_items, const float* const* f, const float* const* cam_norm, float* const* cam_rv,
                              const int* N, const int* h, const int* w, int C, int K, void* const* workspaces,
                              const size_t* workspace_bytes, wsegk_stream_t stream);

/* K1a + K2 forward fused: `cam_logits` [N,C,h,w] are the RAW class logits (the fc8 output); the in-network normalisation
 * (network/resnet38_contrast.py:41-48, what wsegk_cam_norm_fwd computes) is applied inside the operand preparation of the PCM
 * and also written to cam_norm_out [N,C,h,w] (bit-identical to wsegk_cam_norm_fwd; wsegk_pcm_bwd needs it).  Otherwise as
 * wsegk_pcm_fwd. */
int wsegk_pcm_fwd_logits(const float* f, const float* cam_logits, float* cam_norm_out, float* cam_rv, int N, int C, int K,
                         int h, int w, void* workspace, size_t workspace_bytes, wsegk_stream_t stream);

/* K2 backward — autograd of the lines above w.r.t. f only (cam_norm carries no gradient, :41).
 * grad_cam_rv [N,C,h,w] -> grad_f [N,K,h,w].  `workspace` must be the one pcm_fwd filled for the same
 * f / cam_norm (cam_rv is the forward output). */
int wsegk_pcm_bwd(const float* f, const float* cam_norm, const float* cam_rv, const float* grad_cam_rv,
                  float* grad_f, int N, int C, int K, int h, int w, void* workspace, size_t workspace_bytes,
                  wsegk_stream_t stream);

/* The same two operations with the precision mode given PER CALL instead of through the calling thread's default
 * (wsegk_pcm_set_mode) — what a caller with several host threads (torch.nn.DataParallel's parallel_apply,
 * contrast_train.py:108; the thread pool of contrast_infer.py:58-71) should use.
 *   flags: WSEGK_PCM_PRECISE      exact sign mask + split backward factors (the default mode); 0 = "fast"
 *          WSEGK_PCM_DEFER_FIXUP  (forward, precise only) do not run the sign fix-up inside this call: the caller enqueues
 *                                 wsegk_pcm_sign_fixup on the same workspace — e.g. on a side stream, off the critical path —
 *                                 any time after this call and before pcm_bwd.
 *   pcm_fwd_ex: cam_norm_out == NULL: `cam` is the normalised CAM (as wsegk_pcm_fwd); otherwise `cam` holds the raw logits
 *               and cam_norm_out receives the normalised CAM (as wsegk_pcm_fwd_logits).
 *   pcm_bwd_ex: pass the flags of the forward that filled the workspace. */
enum wsegk_pcm_flags { WSEGK_PCM_PRECISE = 1, WSEGK_PCM_DEFER_FIXUP = 2 };
int wsegk_pcm_fwd_ex(const float* f, const float* cam, float* cam_norm_out, float* cam_rv, int N, int C, int K, int h, int w,
                     void* workspace, size_t workspace_bytes, int flags, wsegk_stream_t stream);
int wsegk_pcm_sign_fixup(int N, int C, int K, int h, int w, void* workspace, size_t workspace_bytes, wsegk_stream_t stream);
int wsegk_pcm_bwd_ex(const float* f, const float* cam_norm, const float* cam_rv, const float* grad_cam_rv, float* grad_f,
                     int N, int C, int K, int h, int w, void* workspace, size_t workspace_bytes, int flags,
                     wsegk_stream_t stream);

/* K1b forward — replaces tool/visualization.py:62-67 + contrast_train.py:145-152 / 157-158:
 *   y = relu(relu(p)-min-1e-5)/(max-min+1e-5) per (n,c) plane; bilinear(align_corners) to (oh,ow); * label.
 * p [N,C,H,W], label [N,C] -> out [N,C,oh,ow]; minmax [N,C,4] receives {min, max, argmin, argmax} (the two
 * indices stored as int32 bit patterns) for the backward; tapsum [N,C,oh,ow] (may be NULL when no backward follows)
 * receives, per output, the sum of the bilinear weights of its taps whose normalised input is positive. */
int wsegk_fullres_norm_fwd(const float* p, const float* label, float* out, float* minmax, float* tapsum, int N,
                           int C, int H, int W, int oh, int ow, void* workspace, size_t workspace_bytes,
                           wsegk_stream_t stream);

/* K1b backward — grad_out [N,C,oh,ow] -> grad_p [N,C,H,W], including the gradient routed to each plane's
 * arg-min / arg-max pixel (torch.max/min(dim) backward; first flat index on ties).  `out` and `tapsum` are the
 * forward's outputs: with them the two plane sums of that gradient come from the output-sized tensors, so the
 * full-resolution input is read once. */
int wsegk_fullres_norm_bwd(const float* p, const float* label, const float* minmax, const float* out,
                           const float* tapsum, const float* grad_out, float* grad_p, int N, int C, int H, int W, int oh,
                           int ow, void* workspace, size_t workspace_bytes, wsegk_stream_t stream);

/* K1b on a LAZY full-resolution map: the same result as wsegk_bilinear_resize_fwd(small -> H x W)
 * followed by wsegk_fullres_norm_fwd, bit for bit, without the full-resolution tensor ever existing — i.e.
 * network/resnet38_contrast.py:57-59 (the x8 enlargement at the end of Net.forward) fused into contrast_train.py:145-152.
 * small [N,C,h,w] (the stride-8 cam / cam_rv), label [N,C] -> out [N,C,oh,ow]; minmax and tapsum as above (tapsum required).
 * The backward returns the gradient w.r.t. the SMALL map (what the enlargement's backward would produce from the dense
 * full-resolution gradient): grad_small [N,C,h,w]. */
int wsegk_fullres_norm_up_fwd(const float* small, const float* label, float* out, float* minmax, float* tapsum, int N,
                              int C, int h, int w, int H, int W, int oh, int ow, void* workspace,
                              size_t workspace_bytes, wsegk_stream_t stream);
int wsegk_fullres_norm_up_bwd(const float* small, const float* label, const float* minmax, const float* out,
                              const float* tapsum, const float* grad_out, float* grad_small, int N, int C, int h, int w,
                              int H, int W, int oh, int ow, void* workspace, size_t workspace_bytes,
                              wsegk_stream_t stream);

/* The classification (GAP + multilabel soft margin, contrast_train.py:142,159) and min-pool (:16-25,143) losses on the
 * same lazy full-resolution map: `small` [N,C,h,w] is what Net.forward enlarges to H x W; same arguments otherwise as
 * wsegk_gap_mlsm_fwd/_bwd and wsegk_minpool_fwd/_bwd, gradients on the small map.  m / arg are [N,H*W]. */
int wsegk_gap_mlsm_up_fwd(const float* small, const float* label, float* gap, float* dgap, float* loss, int N, int C, int h,
                          int w, int H, int W, void* workspace, size_t workspace_bytes, wsegk_stream_t stream);
int wsegk_gap_mlsm_up_bwd(const float* dgap, const float* grad_loss, float* grad_small, int N, int C, int h, int w, int H,
                          int W, wsegk_stream_t stream);
int wsegk_minpool_up_fwd(const float* small, const float* label, float* m, int32_t* arg, float* sel, int32_t* idx_cut, int N,
                         int C, int h, int w, int H, int W, void* workspace, size_t workspace_bytes, wsegk_stream_t stream);
int wsegk_minpool_up_bwd(const float* m, const int32_t* arg, const float* sel, const int32_t* idx_cut, const float* label,
                         const float* grad_loss, float* grad_small, int N, int C, int h, int w, int H, int W,
                         void* workspace, size_t workspace_bytes, wsegk_stream_t stream);

/* K1c — replaces contrast_train.py:186-197 / 215-226: relu, per-plane min/max normalisation (no outer relu),
 * channel 0 := bg_threshold, pseudo label = argmax_c(norm_cam*label) (lowest class on ties).
 * cam_rv_down [N,C,g_h,g_w], label [N,C] -> norm_cam [N,C,g_h,g_w], pseudo [N*g_h*g_w] int32. */
int wsegk_seed_labels_fwd(const float* cam_rv_down, const float* label, float bg_threshold, float* norm_cam,
                          int32_t* pseudo, int N, int C, int h, int w, void* workspace,
                          size_t workspace_bytes, wsegk_stream_t stream);

/* Glue — replaces F.interpolate(..., mode='bilinear', align_corners=True): the shrinks of view 1 to the loss grid
 * (contrast_train.py:179-180) and the x8 enlargements at the end of Net.forward (network/resnet38_contrast.py:57-59).
 * x [planes,H,W] -> y [planes,oh,ow], any sizes; and its autograd: grad_y -> grad_x (deterministic gather, no atomics). */
int wsegk_bilinear_resize_fwd(const float* x, float* y, int planes, int H, int W, int oh, int ow, wsegk_stream_t stream);
int wsegk_bilinear_resize_bwd(const float* grad_y, float* grad_x, int planes, int H, int W, int oh, int ow,
                              wsegk_stream_t stream);

/* ---- auxiliary SEAM losses ------------------------------------------------------------------------
 * select_kth: per row of x [rows,len], the k-th largest (largest=1) or smallest (0) value BY VALUE (radix select, no
 * sort).  sel [rows,4] = {threshold t, sum of the k selected values, sum of max(selected,0), number of ties taken};
 * idx_cut [rows]: element i with value v is selected iff v is strictly beyond t, or v == t and i < idx_cut[row]. */
int wsegk_select_kth(const float* x, int rows, int len, int k, int largest, float* sel, int32_t* idx_cut,
                     void* workspace, size_t workspace_bytes, wsegk_stream_t stream);
/* contrast_train.py:142/155 + 159/160: gap = adaptive_avg_pool2d(cam), loss = multilabel_soft_margin(gap[:,1:], label[:,1:]).
 * cam [N,C,H,W], label [N,C] -> gap [N,C], dgap [N,C] (d loss / d gap, kept for the backward), loss [1]. */
int wsegk_gap_mlsm_fwd(const float* cam, const float* label, float* gap, float* dgap, float* loss, int N, int C, int H,
                       int W, void* workspace, size_t workspace_bytes, wsegk_stream_t stream);
int wsegk_gap_mlsm_bwd(const float* dgap, const float* grad_loss, float* grad_cam, int N, int C, int H, int W,
                       wsegk_stream_t stream);
/* contrast_train.py:16-25 on (cam_rv*label)[:,1:]: m [N,H*W] = per-pixel class maximum, arg = its class, then the
 * bottom H*W/4 per image by select_kth; loss = sum(sel[:,2]) / (k*N).  Backward: dense grad_x [N,C,H,W]. */
int wsegk_minpool_fwd(const float* x, const float* label, float* m, int32_t* arg, float* sel, int32_t* idx_cut, int N,
                      int C, int H, int W, void* workspace, size_t workspace_bytes, wsegk_stream_t stream);
int wsegk_minpool_bwd(const float* m, const int32_t* arg, const float* sel, const int32_t* idx_cut, const float* label,
                      const float* grad_loss, float* grad_x, int N, int C, int H, int W, wsegk_stream_t stream);
/* contrast_train.py:163-172 on the four normalised maps [N,C,h,w] (K1b outputs): losses[0] = loss_er,
 * losses[1] = loss_ecr; e1,e2 [N,C*h*w] and sel/cut (see select_kth) are kept for the backward, which writes the
 * gradients of both losses w.r.t. n1, n2 (ER) and rv1, rv2 (ECR). */
int wsegk_er_ecr_fwd(const float* n1, const float* n2, const float* rv1, const float* rv2, float* e1, float* e2,
                     float* sel1, int32_t* cut1, float* sel2, int32_t* cut2, float* losses, int N, int C, int h, int w,
                     void* workspace, size_t workspace_bytes, wsegk_stream_t stream);
int wsegk_er_ecr_bwd(const float* n1, const float* n2, const float* rv1, const float* rv2, const float* sel1,
                     const int32_t* cut1, const float* sel2, const int32_t* cut2, const float* grad_er,
                     const float* grad_ecr, float* gn1, float* gn2, float* grv1, float* grv2, int N, int C, int h, int w,
                     wsegk_stream_t stream);

/* K3 — replaces contrast_train.py:199-209 / 229-241.  Per class, the k largest norm_cam values over the whole
 * batch (value desc, then lowest flat index n*h*w + y*w + x), and the value-weighted feature sums.
 *   proto_topk_local : norm_cam [N,C,h,w], f_proj [N,D,h,w] -> top_val [C,k], top_idx [C,k] (int32, local
 *                      pixel index + pixel_offset), top_fea [C,k,D] (gathered un-normalised features).
 *   proto_finalize   : from R*k candidates per class (concatenated over ranks in rank order) select the global
 *                      top k in canonical order and emit prototypes [C,D] (L2-normalised), sel_idx [C,k]. */
int wsegk_proto_topk_local(const float* norm_cam, const float* f_proj, int pixel_offset, int k, float* top_val,
                           int32_t* top_idx, float* top_fea, int N, int C, int D, int h, int w, void* workspace,
                           size_t workspace_bytes, wsegk_stream_t stream);
int wsegk_proto_finalize(const float* cand_val, const int32_t* cand_idx, const float* cand_fea, int n_cand,
                         int k, float* proto, int32_t* sel_idx, int C, int D, wsegk_stream_t stream);

/* K3 exchange without NCCL: a one-shot gather over NVLink peer memory.  Every rank allocates the same
 * symmetric buffer of wsegk_proto_exchange_bytes(R,C,k,D) bytes, zero-filled once, mapped into every peer
 * (e.g. torch.distributed._symmetric_memory); peer_bases[r] is rank r's buffer as seen from this process.
 *   proto_push          : packs this rank's local top-k (outputs of proto_topk_local) and stores it into its slot of EVERY
 *                         peer's buffer, then releases a per-rank flag carrying the step number (kept in local_state:
 *                         4 uint32, zero-initialised, [view][last step, CTA counter]) — nothing host-side changes between
 *                         steps, so the call is CUDA-graph capturable.
 *   proto_finalize_symm : waits (device-side, bounded) until all R flags of `view` reach this rank's step, then does what
 *                         proto_finalize does on the candidates in place.  Must follow this rank's proto_push of the same
 *                         view on the same stream. */
size_t wsegk_proto_exchange_bytes(int R, int C, int k, int D);
int wsegk_proto_push(const float* top_val, const int32_t* top_idx, const float* top_fea, void* const* peer_bases, int R,
                     int rank, int view, int C, int k, int D, uint32_t* local_state, wsegk_stream_t stream);
int wsegk_proto_finalize_symm(const void* local_base, int R, int view, int k, float* proto, int32_t* sel_idx, int C, int D,
                              const uint32_t* local_state, wsegk_stream_t stream);

/* K4 — replaces contrast_train.py:244-389: the four cross-view NCE terms and the two intra-view NCE terms
 * (semi-hard prototypes = ranks 3..12, hard-pixel mining), forward and backward in one call.
 *   x1,x2   [N,D,h,w]  projected features of the two views on the common loss grid (un-normalised)
 *   proto1/2 [C,D], pl1/pl2 [P] int32 pseudo labels (P = N*h*w)
 *   rand_key1/2 [P] uint32: the "random 50 %" of a class = its pixels with the len//2 smallest
 *            (key, index); pass keys reproducing the reference's Python RNG for bit parity, or any hash.
 *   losses  [8] fp32: {cross_a, cross_b, cross2_a, cross2_b, intra1, intra2, loss_cross_nce, loss_cross_nce2}
 *           then losses[8] = loss_intra_nce, losses[9] = loss_nce (10 floats)
 *   grad_x1/2 [N,D,h,w] = d(loss_nce)/dx * grad_scale_ptr[0] (device scalar; NULL = 1)
 *   mined1/2 [P] int32 (optional, may be NULL): how many times the pixel entered its class's mined set (0..2)
 */
int wsegk_contrast_loss_fwd_bwd(const float* x1, const float* x2, const float* proto1, const float* proto2,
                                const int32_t* pl1, const int32_t* pl2, const uint32_t* rand_key1,
                                const uint32_t* rand_key2, float* losses, float* grad_x1, float* grad_x2,
                                int32_t* mined1, int32_t* mined2, const float* grad_scale, int N, int C, int D,
                                int h, int w, void* workspace, size_t workspace_bytes, wsegk_stream_t stream);

/* K4 with GLOBAL hard-pixel mining under data parallelism: as wsegk_contrast_loss_fwd_bwd on this
 * rank's shard, but the per-class order statistics of the mining (contrast_train.py:302-322) are taken over the pixels of ALL R
 * ranks, as the reference's single-process loss does.  Every rank pushes (sim, key, pseudo label) of its pixels into every
 * peer's symmetric buffer (wsegk_mine_exchange_bytes; peer_bases[r] = rank r's buffer mapped here, local_base = mine;
 * local_state: 2 zero-initialised uint32 owned by the library) and ranks its own pixels against all of them in the canonical
 * order (value, global pixel = rank*P + p).  Reported intra losses are R x this rank's share: their mean over ranks — what
 * DDP does to the gradients — is the single-process value.  Up to 16 384 loss pixels per rank the local pixels are ranked
 * pairwise against the gathered batch; above that every rank runs the multi-CTA select over all R*P gathered pixels (three
 * threshold keys per class, the same on every rank) and membership is two comparisons per pixel. */
size_t wsegk_mine_exchange_bytes(int R, int N, int h, int w);
int wsegk_contrast_loss_fwd_bwd_global(const float* x1, const float* x2, const float* proto1, const float* proto2,
                                       const int32_t* pl1, const int32_t* pl2, const uint32_t* rand_key1,
                                       const uint32_t* rand_key2, float* losses, float* grad_x1, float* grad_x2,
                                       int32_t* mined1, int32_t* mined2, const float* grad_scale, int N, int C, int D,
                                       int h, int w, void* workspace, size_t workspace_bytes, void* const* peer_bases,
                                       const void* local_base, int R, int rank, uint32_t* local_state,
                                       wsegk_stream_t stream);

/* Mining keys generated on the device (stands in for the random.sample draws of contrast_train.py:316,370):
 * keys[i] = Philox4x32-10(counter = (index0 + i, step, stream_id), key = seed)[0].  Counter-based: no state, no host work,
 * and the same key for a pixel however the batch is sharded (index0 = the shard's first global pixel). */
int wsegk_mining_keys(uint32_t* keys, int count, unsigned long long index0, unsigned long long seed,
                      unsigned long long step, unsigned int stream_id, wsegk_stream_t stream);

/* Fused multi-tensor SGD step (torch.optim.SGD semantics: d = grad + wd*p; buf = first ? d : momentum*buf + (1-dampening)*d;
 * p -= lr*buf) for the reference's PolyOptimizer (tool/torchutils.py:11-33; the step after the hot path).
 * The four tables are HOST arrays of n_tensors device pointers / element counts; tensors are contiguous fp32. */
int wsegk_sgd_step(float* const* params, const float* const* grads, float* const* momentum_bufs, const long long* numels,
                   int n_tensors, float lr, float weight_decay, float momentum, float dampening, int first_step,
                   wsegk_stream_t stream);

/* Diagnostics for bench.py / tests (not part of the reference-facing surface).
 *   launch_count  : kernels this library has enqueued in this process so far.
 *   pcm_profile   : when enabled (per host thread), pcm_fwd / pcm_bwd record CUDA events on `stream` around their
 *                   tcgen05 main kernel; pcm_main_ms synchronises on them and returns the LAST durations in ms
 *                   (negative if that kernel has not run with profiling on). */
unsigned long long wsegk_debug_launch_count(void);
int wsegk_debug_pcm_profile(int enable);
int wsegk_debug_pcm_main_ms(float* fwd_ms, float* bwd_ms);
int wsegk_debug_pcm_fixup_ms(float* fix_ms);   /* the same for the sign fix-up kernel of the precise forward */
/* Timeline aid: when non-null (per host thread), CTA (0,0,0) of the next pcm_fwd main kernels writes clock64() stamps
 * into this DEVICE buffer of 768 int64: [role 0 producer | 1 MMA issuer | 2 first TMEM warp][stage < 64][4 events]. */
int wsegk_debug_pcm_trace(long long* device_buffer);
/* diagnostics of the NVLink prototype exchange: device buffer of 2 x 17 uint64 globaltimer stamps per view — [0] the rank
   started waiting for its peers' pushes, [1 + r] it saw rank r's flag; NULL switches it off (thread-local) */
int wsegk_debug_exchange_stamps(unsigned long long* device_buffer);
/* diagnostics: *dst = %globaltimer (ns) when `stream` reaches this point */
int wsegk_debug_stamp(unsigned long long* dst, wsegk_stream_t stream);
/* the same for the backward main kernel (768 int64) */
int wsegk_debug_pcm_trace_bwd(long long* device_buffer);
/* Test aid (host only, no GPU needed): the (unit, stage, segment, partial slot) quadruples CTA pair g of the persistent PCM
 * kernels runs for N images on an h x w grid; returns their number (0 for g beyond the grid). */
int wsegk_debug_pcm_schedule(int N, int h, int w, int g, int* out, int capacity);
/* Test aid, per host thread: when on, wsegk_proto_topk_local and wsegk_contrast_loss_fwd_bwd take their large-plane paths
 * (multi-CTA radix select, csrc/segselect.cuh) whatever the size, so both paths can be compared on the same inputs.
 * By default K3 switches above 8192 pixels and the mining of K4 above 16384. */
int wsegk_debug_force_radix(int on);
/* Test aid, per host thread: which part of wsegk_contrast_loss_fwd_bwd_global the next calls run — bit 0: everything up to
 * and including the push of this rank's (sim, key, label) to its peers; bit 1: the rest (wait for the peers, mining, gradient).
 * Default 3.  With 1 then 2 (same arguments, same workspace) ONE process can drive several "ranks" of one GPU on one stream:
 * all pushes first, then all second halves — no kernel ever spins on a flag that a later launch of the same queue sets. */
int wsegk_debug_mine_phases(int mask);

/* PCM precision mode used by the entry points without a `flags` argument, per HOST THREAD (default 1):
 *   1 "precise": S from single FP16 products; the entries with |S| below the bound of the operand-rounding error (~0.2 %)
 *                are re-evaluated in fp32 (pcm_sign_fixup_kernel), so the sign mask of S that the backward consumes agrees
 *                with an FP32 evaluation; backward factors split in BF16 (hi + lo).
 *   0 "fast"   : no re-evaluation, single BF16 factor products (cam_rv identical; gradient terms with |S| < ~1e-4 may flip). */
int wsegk_pcm_set_mode(int mode);
int wsegk_pcm_get_mode(void);
/* CTA pairs (1 .. 74; 0 = the default) the persistent PCM forward / backward main kernels of the calling THREAD run on.
 * The defaults (64 / 52) are tuned for the whole training step, where the SMs a PCM kernel leaves free run the other chains'
 * kernels; a PCM call on an otherwise idle GPU is fastest with 72 / 72.  The workspace layout depends on the setting: do not
 * change it between wsegk_workspace_bytes, a forward and its backward. */
int wsegk_pcm_set_pairs(int fwd_pairs, int bwd_pairs);

/* Test aid: byte offsets of the PCM workspace sections, in the order
 * {fa, fh32, fb, nzmask, camb, inv_norm, dsum, part, mask, ua, vb, dfhat, bound, total_bytes, nsplit, PJ, HWpj,
 *  wb, NS, HWpi}.  Returns how many were written. */
int wsegk_debug_pcm_layout(int N, int C, int K, int h, int w, size_t* out, int n_out);

#ifdef __cplusplus
}
#endif
#endif /* WSEGK_H_ */
