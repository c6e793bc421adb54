/* libunetb200 — C ABI of the B200-native U-Net(ResNet-34) hot path.
 *
 * This is the drop-in boundary for the one path BASELINE.json names: what the reference reaches through
 *   segmentation_models_pytorch.Unet("resnet34", in_channels=3, classes=1)(x)
 *     (/root/reference/train.py:372-378,436 ; infer_pth_gui.py:31-33,50-51 ; ui_infer_rectangle.py:496-499,556-559 ;
 *      ui_infer_quadrilateral.py:638-641,704-707)
 *   nn.BCEWithLogitsLoss() + smp.losses.DiceLoss(mode="binary")      (/root/reference/train.py:600-601,438)
 *   loss.backward() / torch.optim.AdamW(...).step()                   (/root/reference/train.py:441-449,606)
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; every pointer named *_dev is a device pointer owned by the caller
 *     (PyTorch) and valid for the duration of the stream-ordered call; the library never frees or retains it.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are asynchronous.
 *   - every function returns 0 on success, non-zero on error; unetb200_last_error() gives the message.
 *   - there is no CPU or cuDNN fallback: a non-sm_100 device, a bad shape or a missing weight load is an error.
 *   - one ctx per (process, device); calls on one ctx are not re-entrant.
 *
 * Tensor table: the library defines the canonical order (== state_dict order of smp.Unet("resnet34"), 278 entries,
 * SURVEY.md section 8b).  kind 0 = fp32 parameter (offset into the flat parameter array, 24,436,369 floats),
 * kind 1 = fp32 buffer running_mean/var (offset into the flat buffer array), kind 2 = int64 num_batches_tracked
 * (offset = index into the counter array).
 */
#ifndef UNETB200_H
#define UNETB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct unetb200_ctx unetb200_ctx;

/* ---- life cycle ------------------------------------------------------------------------------------------------ */
/* replaces: smp.Unet(...).to(device)  (train.py:595, infer_pth_gui.py:43).  H, W: input size, multiples of 32. */
int unetb200_create(unetb200_ctx** out, int device, int max_batch, int H, int W);
void unetb200_destroy(unetb200_ctx* ctx);
/* ctx may be NULL: returns the message of the last failed unetb200_create on this thread. */
const char* unetb200_last_error(unetb200_ctx* ctx);
/* device-side pipeline watchdog flag (0 = healthy); synchronises the device. */
int unetb200_check_device_error(unetb200_ctx* ctx, int* flag_out);

/* ---- tensor table (state_dict layout) ---------------------------------------------------------------------------- */
int unetb200_num_tensors(void);
int unetb200_tensor_info(int index, char* name_out, int name_cap, int* ndim_out, int shape_out[4],
                         long long* offset_out, int* kind_out);
long long unetb200_num_params(void);   /* 24,436,369 */
long long unetb200_num_buffers(void);  /* 19,008 fp32 running statistics */
int unetb200_num_counters(void);       /* 46 int64 num_batches_tracked */

/* ---- weights ----------------------------------------------------------------------------------------------------- */
/* replaces: model.load_state_dict(sd) (infer_pth_gui.py:42) / the implicit use of current parameters by forward.
 * Re-packs the fp32 master tensors into the bf16 K-major operand caches and folds eval-mode BatchNorm. */
int unetb200_load_weights(unetb200_ctx* ctx, const float* params_dev, const float* buffers_dev, void* stream);
/* same; skip_bn_fold != 0 leaves the folded eval-mode BatchNorm constants untouched (a training step uses batch
 * statistics, so the fold is only needed again before the next eval-mode forward). */
int unetb200_load_weights_ex(unetb200_ctx* ctx, const float* params_dev, const float* buffers_dev, int skip_bn_fold,
                             void* stream);

/* ---- inference: model.eval(); torch.sigmoid(model(x)) >= t  (infer_pth_gui.py:50-52) ---------------------------- */
/* x_dev: fp32 NCHW [N,3,H,W].  Any of logits/prob/mask may be NULL (at least one must not be):
 *   logits_dev fp32 [N,1,H,W]; prob_dev fp32 sigmoid(logits); mask_dev uint8 {0,255} = prob >= thresh. */
int unetb200_forward_infer(unetb200_ctx* ctx, const float* x_dev, float* logits_dev, float* prob_dev,
                           uint8_t* mask_dev, float thresh, int N, void* stream);
/* Same through HOST buffers (pinned or pageable): H2D of x, forward, D2H of the requested outputs, synchronous.
 * This is the end-to-end call a non-PyTorch host (the reference GUIs' numpy path) would bind. */
int unetb200_infer_host(unetb200_ctx* ctx, const float* x_host, float* logits_host, float* prob_host,
                        uint8_t* mask_host, float thresh, int N);
/* Pipelined form of the same call: submit returns as soon as the work is enqueued; wait blocks until the outputs of
 * that slot are in the host buffers.  Two slots (0 / 1): while slot k computes, the H2D copy of the other slot's next
 * request and the D2H copy of its previous result run on the copy engines (three streams inside the library), which
 * is how a caller streaming frames keeps PCIe traffic off the critical path.  Host buffers should be pinned; they
 * must stay valid until the matching wait.  A slot must be waited on before it is submitted again. */
int unetb200_infer_host_submit(unetb200_ctx* ctx, int slot, const float* x_host, float* logits_host,
                               float* prob_host, uint8_t* mask_host, float thresh, int N);
int unetb200_infer_host_wait(unetb200_ctx* ctx, int slot);
/* Camera-frame input: img uint8 HWC [N,H,W,3] (bgr != 0: B,G,R order as cv2.imread returns it).  The reference's host
 * pre-processing — BGR->RGB, /255, (x - mean) / std (infer_pth_gui.py:46-48, ui_infer_rectangle.py:530-533) — runs
 * inside the input-pack kernel; mean3 / std3 are the per-RGB-channel constants (NULL: 0 / 1).  Letterboxing stays on
 * the host (it is an OpenCV resize).  4x fewer bytes cross the bus than with the fp32 tensor. */
int unetb200_infer_host_u8(unetb200_ctx* ctx, const uint8_t* img_host, int bgr, const float* mean3, const float* std3,
                           float* logits_host, float* prob_host, uint8_t* mask_host, float thresh, int N);
int unetb200_infer_host_u8_submit(unetb200_ctx* ctx, int slot, const uint8_t* img_host, int bgr, const float* mean3,
                                  const float* std3, float* logits_host, float* prob_host, uint8_t* mask_host,
                                  float thresh, int N);
int unetb200_forward_infer_u8(unetb200_ctx* ctx, const uint8_t* img_dev, int bgr, const float* mean3,
                              const float* std3, float* logits_dev, float* prob_dev, uint8_t* mask_dev, float thresh,
                              int N, void* stream);
/* number of kernel launches one forward_infer call issues at batch N (after the plan for N exists) */
int unetb200_infer_launch_count(unetb200_ctx* ctx, int N);

/* Test hooks: the activations the last forward_infer at batch N left in the library's arena, by index: NHWC bf16
 * [n,h,w,c].  Names: "xp" (packed input [N,H,W+8,4]), "<conv weight key>/out" (the tensor that launch wrote: after the
 * folded BatchNorm (+residual) (+ReLU)), "<decoder conv1 key>/up" (scaled partial over the up-sampled channels),
 * "encoder.maxpool/out".  The layer-local parity tests re-derive every launch's output from ITS OWN inputs. */
int unetb200_infer_debug_count(unetb200_ctx* ctx, int N);
int unetb200_infer_debug_info(unetb200_ctx* ctx, int N, int index, char* name_out, int name_cap, int shape_out[4]);
int unetb200_infer_debug_copy(unetb200_ctx* ctx, int N, int index, void* dst_dev, long long cap_bytes, void* stream);

/* One forward with a CUDA event recorded between consecutive launches on `stream`: ms_out[i] = device time of launch i
 * (0 = input pack, last = seg head), is_igemm_out[i] = 1 for implicit-GEMM conv launches.  Synchronises the stream. */
int unetb200_profile_infer(unetb200_ctx* ctx, const float* x_dev, float* logits_dev, int N, void* stream,
                           float* ms_out, int* is_igemm_out, int cap, int* n_out);
int unetb200_profile_name(unetb200_ctx* ctx, int N, int index, char* name_out, int name_cap);

/* ---- per-kernel entry points for unit parity tests --------------------------------------------------------------- */
/* conv + folded scale/shift (+residual) (+ReLU) on NHWC bf16: w_dev fp32 OIHW [cout,cin,k,k], k in {1,3},
 * stride in {1,2}, padding k/2.  scale/shift/residual may be NULL.  stats_dev (optional) fp32 [cout][2] receives
 * per-channel sum / sum-of-squares of the bf16 output. */
int unetb200_conv_nhwc(unetb200_ctx* ctx, const void* in_bf16_dev, const float* w_dev, const float* scale_dev,
                       const float* shift_dev, const void* residual_bf16_dev, int relu, void* out_bf16_dev,
                       float* stats_dev, int N, int H, int W, int cin, int cout, int k, int stride, void* stream);

/* ---- training step (/root/reference/train.py:428-449) --------------------------------------------------------------- */
/* model.train(); logits = model(x)  (train.py:413,436).  BatchNorm uses batch statistics; running_mean / running_var
 * (buffers_dev, fp32 flat) and num_batches_tracked (counters_dev, int64[46]) are updated in place like nn.BatchNorm2d
 * (momentum 0.1, unbiased running variance).  Activations needed by the backward stay in the library's arena until the
 * next train_forward.  grads_dev is the flat fp32 gradient array (same layout as params_dev) train_backward fills.
 * unetb200_load_weights must have been called since the parameters last changed. */
int unetb200_train_forward(unetb200_ctx* ctx, const float* x_dev, float* logits_dev, const float* params_dev,
                           float* buffers_dev, long long* counters_dev, float* grads_dev, int N, void* stream);
/* Same with uint8 HWC frames [N,H,W,3] as the input (what cv2.imread / the letterbox of train.py:70-75 produce): BGR->RGB,
 * /255 and (x - mean) / std of train.py:108-112 run inside the input pack, so a data loader ships 0.75 MB instead of
 * 3 MB per 512x512 image to the GPU. */
int unetb200_train_forward_u8(unetb200_ctx* ctx, const uint8_t* img_dev, int bgr, const float* mean3, const float* std3,
                              float* logits_dev, const float* params_dev, float* buffers_dev, long long* counters_dev,
                              float* grads_dev, int N, void* stream);
/* loss.backward() through the network (train.py:443,448) given dL/dlogits fp32 [N,1,H,W].  The backward is cut into
 * 4 stages whose parameter gradients are complete when the stage ends (0 = head + decoder, 1 = encoder.layer4,
 * 2 = layer3, 3 = layer2 + layer1 + stem) so that a data-parallel caller can all-reduce bucket k while stage k+1 runs.
 * Runs stages stage_first..stage_last; stage 0 first OVERWRITES the whole gradient array (no accumulation). */
int unetb200_train_backward(unetb200_ctx* ctx, const float* dlogits_dev, int N, int stage_first, int stage_last,
                            void* stream);
/* element range [begin, end) of the flat parameter / gradient array that backward stage `stage` completes */
int unetb200_grad_bucket_range(int stage, long long* begin_out, long long* end_out);
/* kernel launches of one train_forward / one full train_backward at batch N (after the plan exists) */
int unetb200_train_launch_count(unetb200_ctx* ctx, int N, int* fwd_out, int* bwd_out);

/* Test hooks: the library's internal training tensors of the last step at batch N (saved activations z / a, their
 * gradients dz / dA, batch mean / invstd, ...), by index; NHWC, bf16 (is_bf16 = 1) or fp32.  Used by the layer-local
 * parity tests, which re-derive every backward kernel's output from ITS OWN inputs with PyTorch. */
int unetb200_train_debug_count(unetb200_ctx* ctx, int N);
int unetb200_train_debug_info(unetb200_ctx* ctx, int N, int index, char* name_out, int name_cap, int shape_out[4],
                              int* is_bf16_out);
int unetb200_train_debug_copy(unetb200_ctx* ctx, int N, int index, void* dst_dev, long long cap_bytes, void* stream);

/* Per-launch timing of train_forward / train_backward: while enabled a CUDA event is recorded on the launch stream
 * before every kernel launch; profile_dump synchronises and writes "launch,kind,layer,ms" rows (CSV) to `path`. */
int unetb200_profile_enable(unetb200_ctx* ctx, int on);
int unetb200_profile_dump(unetb200_ctx* ctx, const char* path);

/* nn.BCEWithLogitsLoss()(logits, y) + smp.losses.DiceLoss("binary")(logits, y)  (train.py:438,600-601), one pass:
 * result_dev[0..2] = bce, dice, bce + dice; result_dev[3..5] = sum(p*y), sum(p), sum(y) (kept for the backward);
 * result_dev must hold 8 floats.  n = N*H*W elements; eps = DiceLoss eps (1e-7).  scratch_dev: caller-owned device
 * scratch of unetb200_loss_scratch_floats() floats.  No ctx: errors are reported through unetb200_last_error(NULL). */
int unetb200_loss_scratch_floats(void);
int unetb200_loss_bce_dice_forward(const float* logits_dev, const float* target_dev, long long n, float eps,
                                   float* scratch_dev, float* result_dev, void* stream);
/* dlogits = gscale * (g_bce * dBCE/dlogits + g_dice * dDice/dlogits); g_*_dev are device scalars (NULL = 0). */
int unetb200_loss_bce_dice_backward(const float* logits_dev, const float* target_dev, const float* result_dev,
                                    const float* g_bce_dev, const float* g_dice_dev, float gscale, float eps,
                                    float* dlogits_dev, long long n, void* stream);

/* ---- validation metrics (/root/reference/train.py:230-281 dice_coef / iou_coef, called at train.py:518-522) ----------- */
/* pred_dev: fp32 [N,1,H,W] probabilities (thresh 0.5) or logits (thresh 0: sigmoid(x) > 0.5 <=> x > 0); target_dev: fp32
 * {0,1} masks of the same shape; hw = H*W.  out2_dev[0] = batch mean of the per-image Dice, out2_dev[1] = of the IoU,
 * with the reference's eps placement.  scratch_dev: unetb200_seg_metrics_scratch_floats(N) floats.  Deterministic. */
int unetb200_seg_metrics_scratch_floats(int N);
int unetb200_seg_metrics(const float* pred_dev, const float* target_dev, int N, long long hw, float thresh, float eps,
                         float* scratch_dev, float* out2_dev, void* stream);

/* torch.optim.AdamW(...).step() (train.py:606,444,449) fused over the flat arrays: decoupled weight decay on every
 * element, bias-corrected moments (step >= 1).  grad_scale multiplies the gradient first (1/world_size, 1/loss_scale);
 * zero_grad != 0 clears grads_dev afterwards (optimizer.zero_grad, train.py:428). */
int unetb200_adamw_step(unetb200_ctx* ctx, float* params_dev, float* grads_dev, float* exp_avg_dev,
                        float* exp_avg_sq_dev, long long n, float lr, float beta1, float beta2, float eps,
                        float weight_decay, long long step, float grad_scale, int zero_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETB200_H */
