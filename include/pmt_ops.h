/*
 * pmt_ops.h -- C ABI of libpmt_ops.so: B200 (sm_100a) kernels for the stereo cost-volume hot path of
 * cuevhv/PMT_learning_for_semantic_segmentation_and_disparity.
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; no torch types.  Every pointer is a DEVICE pointer to a dense
 *     row-major fp32 tensor unless its name ends in `_host`.
 *   - the caller owns every buffer; the library never allocates, frees or retains device memory
 *     (exception: the `_host` convenience entry point keeps a grow-only per-device scratch + streams).
 *   - `stream` is a cudaStream_t passed as void*; kernels are only enqueued, never synchronised
 *     (the `_host` entry points synchronise before returning because they hand back host data).
 *   - return value 0 = success; non-zero = error, message via pmt_last_error() (thread-local).
 *   - there is no CPU fallback anywhere in this library.
 *
 * Each entry point cites the reference interface it replaces (file:line in the reference repo).
 */
#ifndef PMT_OPS_H_
#define PMT_OPS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMT_OK 0
#define PMT_ERR_INVALID 1   /* bad shape / argument */
#define PMT_ERR_CUDA 2      /* CUDA runtime / driver error */
#define PMT_ERR_UNSUPPORTED 3

/* library version (major*10000 + minor*100 + patch) and last error of the calling thread */
int pmt_version(void);
const char* pmt_last_error(void);
/* 1 if device `dev` can run this library (compute capability 10.x), else 0 */
int pmt_device_supported(int dev);

/* ---------------------------------------------------------------------------------------------
 * a1. spatial_correlation_sampler backend (third-party pybind module
 *     `spatial_correlation_sampler_backend.forward/backward`), as constructed at
 *     models/dsnet_t2.py:129-133,425,847-851,1078-1087; models/dsnet_t2_warp.py:197,506,615,742,877;
 *     models/torch_dsnet.py:133-138; models_deeplab_mod/net.py:99-103  (kernel_size=1, stride=1,
 *     padding=0, dilation=1 at every call site).
 *
 *   out[n,ph,pw,h,w] = sum_c in1[n,c,h,w] * in2[n,c,h+sh,w+sw]   (terms outside the image skipped)
 *   sh = (ph-(patchH-1)/2)*dilpH, sw = (pw-(patchW-1)/2)*dilpW;  out is (B,patchH,patchW,H,W).
 *
 * pmt_corr1d_{fwd,bwd}_f32: the 1 x P horizontal patch (the hot path; `-corrType 1dcorr`), fp32-accurate
 *   results from the fastest engine that fits: tensor cores with the 3xTF32 split (W%4==0, 16-byte
 *   pointers, dilp==1; forward P<=193, backward P<=256; any C -- channel blocks of 128 in the backward) ->
 *   CUDA-core TMA-tiled kernels -> generic kernels.  Backward is a deterministic gather on every engine (bit-reproducible run to run).
 * pmt_corr1d_{fwd,bwd}_simt_f32: force the CUDA-core engines (fp32 FFMA).
 * pmt_corr1d_{fwd,bwd}_tc_f32: force the tensor-core engine (see below).
 * pmt_corr_*: any (patchH, patchW, dilation_patch) -- generic CUDA kernels (2-D 17x17 patches of
 *   `-corrType 2dcorr`, the (1,21) dilation_patch=4 sampler of torch_dsnet.py:133-138).
 * ------------------------------------------------------------------------------------------- */
int pmt_corr1d_fwd_f32(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                       int P, int dilp, void* stream);
int pmt_corr1d_bwd_f32(const float* in1, const float* in2, const float* gout, float* gin1,
                       float* gin2, int B, int C, int H, int W, int P, int dilp, void* stream);
int pmt_corr1d_fwd_simt_f32(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                            int P, int dilp, void* stream);
int pmt_corr1d_bwd_simt_f32(const float* in1, const float* in2, const float* gout, float* gin1,
                            float* gin2, int B, int C, int H, int W, int P, int dilp, void* stream);
/* Tensor-core forward/backward of the 1 x P correlation (tcgen05.mma kind::tf32, accumulator in TMEM).
 *   passes = 1: plain TF32 inputs (reduced-precision variant, ~1e-3 relative);
 *   passes = 3: 3xTF32 split (hi*hi + hi*lo + lo*hi), fp32-class accuracy (<= 1e-5 relative).
 * Returns PMT_ERR_UNSUPPORTED when the shape/alignment does not fit (no fallback inside). */
int pmt_corr1d_fwd_tc_f32(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                          int P, int dilp, int passes, void* stream);
int pmt_corr1d_bwd_tc_f32(const float* in1, const float* in2, const float* gout, float* gin1,
                          float* gin2, int B, int C, int H, int W, int P, int dilp, int passes,
                          void* stream);
int pmt_corr_fwd_f32(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                     int patchH, int patchW, int dilpH, int dilpW, void* stream);
int pmt_corr_bwd_f32(const float* in1, const float* in2, const float* gout, float* gin1,
                     float* gin2, int B, int C, int H, int W, int patchH, int patchW, int dilpH,
                     int dilpW, void* stream);
/* f2 (next row): the 1 x P correlation fused with its epilogue
 *     y = squeeze(correlation_sampler(a, b), 1); y = corrConv2d(y)        models/dsnet_t2.py:1187-1197, :879-888
 *     corrConv2d = conv2dSame(P, O, 1, 'same') [bias=False] + ReLU        models/dsnet_t2.py:852, dsnet_t2_warp.py:664-671
 *   z[n,o,h,w] = relu(sum_p weight[o,p] * sum_c in1[n,c,h,w]*in2[n,c,h,w+s_p]);  weight is (O,P) (= Conv2d weight
 *   (O,P,1,1)); corr_save (B,P,H,W) receives the correlation slab (needed by the backward for gweight).
 *   Backward: gin1, gin2, gweight (O,P) from gz = d/dz; workspace holds B*H*O*P floats (per-row partials of gweight,
 *   reduced in fixed order: deterministic).  One CTA per image row; covers the production shape of the training
 *   configurations (C=352, 32x64, P=17, O=128): pmt_corr1d_conv_relu_supported() tells (P==17, 16<=W<=128, W%4==0,
 *   O<=256); other shapes return PMT_ERR_UNSUPPORTED (callers keep the unfused sampler -> conv -> relu sequence). */
int pmt_corr1d_conv_relu_supported(int C, int H, int W, int P, int O);
int pmt_corr1d_conv_relu_fwd_f32(const float* in1, const float* in2, const float* weight, float* z, float* corr_save,
                                 int B, int C, int H, int W, int P, int O, void* stream);
int pmt_corr1d_conv_relu_bwd_f32(const float* in1, const float* in2, const float* weight, const float* z,
                                 const float* corr_save, const float* gz, float* gin1, float* gin2, float* gweight,
                                 float* workspace, int B, int C, int H, int W, int P, int O, void* stream);
/* which engine pmt_corr1d_{fwd,bwd}_f32 would use: 2 = tensor core (3xTF32), 1 = CUDA-core tiled, 0 = generic */
int pmt_corr1d_uses_fast_path(const void* in1, const void* in2, const void* out_or_gout, int C,
                              int H, int W, int P, int dilp);

/* ---------------------------------------------------------------------------------------------
 * a2. PSMNet concat cost volume -- replaces the slice-assign loop of
 *     models_psmnet/stackhourglass.py:110-119 (and matchshifted, models_psmnet/submodule.py:45-54,
 *     which is the single-plane case D=1 with plane index `first_disp`).
 *   cost[b, c,   i, h, w] = ref[b,c,h,w]          if w >= d else 0     d = first_disp + i
 *   cost[b, C+c, i, h, w] = tgt[b,c,h,w-d]        if w >= d else 0     i in [0, D)
 *   Every element of `cost` (B,2C,D,H,W) is written (no reliance on pre-zeroed memory).
 *   Backward: gref[b,c,h,w] = sum_i gcost[b,c,i,h,w][w>=d];  gtgt[b,c,h,w'] = sum_i gcost[b,C+c,i,h,w'+d]
 * ------------------------------------------------------------------------------------------- */
int pmt_concat_volume_fwd_f32(const float* ref, const float* tgt, float* cost, int B, int C, int D,
                              int H, int W, int first_disp, void* stream);
int pmt_concat_volume_bwd_f32(const float* gcost, float* gref, float* gtgt, int B, int C, int D,
                              int H, int W, int first_disp, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a3. disparityregression (models_psmnet/submodule.py:56-64) and the fused
 *     F.softmax(dim=1) + disparityregression pair (models_psmnet/stackhourglass.py:142-155).
 *   dispreg:    out[b,h,w] = sum_d x[b,d,h,w] * d ;            gx[b,d,h,w] = d * gout[b,h,w]
 *   softargmin: out = sum_d d * softmax(cost)[d]  (single pass, online softmax);
 *               `lse` (B,H,W) receives max + log(sum exp) for the backward (may be NULL).
 *               gcost[b,d,h,w] = gout * p_d * (d - out),  p_d = exp(cost - lse)
 * ------------------------------------------------------------------------------------------- */
int pmt_dispreg_fwd_f32(const float* x, float* out, int B, int D, int H, int W, void* stream);
int pmt_dispreg_bwd_f32(const float* gout, float* gx, int B, int D, int H, int W, void* stream);
int pmt_softargmin_fwd_f32(const float* cost, float* out, float* lse, int B, int D, int H, int W,
                           void* stream);
int pmt_softargmin_bwd_f32(const float* cost, const float* out, const float* lse, const float* gout,
                           float* gcost, int B, int D, int H, int W, void* stream);

/* f1 (next row): F.upsample(cost3, [D,H,W], mode='trilinear') fused into the soft-argmin --
 * models_psmnet/stackhourglass.py:149-155 (and :138-147 for pred1/pred2).  `lowres` is (B,Dq,Hq,Wq) (the squeezed
 * (B,1,Dq,Hq,Wq) logits); the (B,D,H,W) upsampled volume is never materialised.  align_corners=False semantics of
 * ATen (scale = in/out, src = scale*(dst+0.5)-0.5 clamped at 0).  `lse` (B,H,W) may be NULL in the forward.
 * Backward: glowres = d(sum gout*out)/d lowres without the volume either: a first kernel writes, per output pixel,
 * the gradient w.r.t. the Dq bilinearly sampled source planes into `workspace` (B*Dq*H*W floats, caller-allocated),
 * a second one applies the adjoint of the spatial interpolation (deterministic gather).  `out` and `lse` are the
 * forward's results.  pmt_upsample_softargmin_bwd_supported() = 1 when the shape fits the fused kernels (shared-memory
 * footprint, interpolation windows <= 40); otherwise the caller must differentiate the unfused sequence. */
int pmt_upsample_softargmin_fwd_f32(const float* lowres, float* out, float* lse, int B, int Dq, int Hq,
                                    int Wq, int D, int H, int W, void* stream);
int pmt_upsample_softargmin_bwd_supported(int B, int Dq, int Hq, int Wq, int D, int H, int W);
int pmt_upsample_softargmin_bwd_f32(const float* lowres, const float* out, const float* lse, const float* gout,
                                    float* workspace, float* glowres, int B, int Dq, int Hq, int Wq, int D, int H,
                                    int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4. apply_disparity(input_images, x_offset, wrap_mode='edge') -- models/torch_dsnet.py:10-86.
 *   x = clamp(w + off[n,0,h,w], 0, W-1); x0 = floor(x); x1 = min(x0+1, W-1)
 *   out[n,c,h,w] = (x1-x)*img[n,c,h,x0] + (x-x0)*img[n,c,h,x1]       (fp32 steps as the reference)
 *   `out` is written in the reference's storage order [C][N][H][W] when out_cnhw != 0 (the
 *   reference returns that buffer permuted to (N,C,H,W), torch_dsnet.py:84), else [N][C][H][W].
 *   The flat gather index is formed in fp32 exactly as torch_dsnet.py:59-70 does (only matters when
 *   N*H*W >= 2^24).  Backward: every element of gimg is WRITTEN by the call (no zero fill by the caller).
 *   While N*H*W < 2^24 and W <= 1024 (pmt_warp1d_rows_supported() == 1) the taps of a pixel stay in its image
 *   row and the backward is a deterministic per-row gather (CSR of the taps built once per row in shared
 *   memory): bit-reproducible, no atomics.  Otherwise it zeroes gimg and scatters with fp32 atomics like the
 *   reference's gather backward.
 *   goff[n,0,h,w] = sum_c gout*(img[x1]-img[x0]) where 0 <= w+off <= W-1, else 0.
 *   `gout` uses the same storage order flag as `out`.  gimg or goff may be NULL (not both).
 * ------------------------------------------------------------------------------------------- */
int pmt_warp1d_fwd_f32(const float* img, const float* off, float* out, int N, int C, int H, int W,
                       int out_cnhw, void* stream);
int pmt_warp1d_bwd_f32(const float* img, const float* off, const float* gout, float* gimg,
                       float* goff, int N, int C, int H, int W, int gout_cnhw, void* stream);
int pmt_warp1d_rows_supported(int N, int H, int W);

/* f4 (next row): the warp fused with its two consumers.  All tensors (N,C,H,W) / (N,1,H,W)
 * dense NCHW.
 *   blend -- models/dsnet_t2_warp.py:697-698:  warped = apply_disparity(img, off);
 *            out = (1 - att) * seg + att * warped  (the reference's three rounded fp32 steps: bit-identical).
 *            `warped` (may be NULL) receives the warped tensor the model returns as well.
 *            backward: gout = d/d out, gwarped = d/d warped (may be NULL);
 *              gseg = (1-att)*gout;  gatt = sum_c gout*(warped - seg);  gimg/goff = warp backward of att*gout + gwarped.
 *   photo-consistency MSE -- torch_implementation.py:314-317 with warped_right = apply_disparity(right, -disp)
 *            [* (disp > 0) when mask_positive_disp, models/dsnet_t2_warp.py:811]:
 *            loss = mean((warped*mask - left)^2), reduced in a fixed order (bit-reproducible).  `workspace` holds
 *            pmt_warp1d_mse_workspace() doubles.  backward: gloss = device scalar d/d loss (NULL = 1);
 *            gleft = -2*gloss/numel*(warped*mask - left); gimg/goff = warp backward of mask * (-gleft).
 * The backward entry points need pmt_warp1d_rows_supported(N,H,W) == 1. */
int pmt_warp1d_blend_fwd_f32(const float* img, const float* off, const float* att, const float* seg, float* out,
                             float* warped, int N, int C, int H, int W, void* stream);
int pmt_warp1d_blend_bwd_f32(const float* img, const float* off, const float* att, const float* seg, const float* gout,
                             const float* gwarped, float* gimg, float* goff, float* gatt, float* gseg, int N, int C,
                             int H, int W, void* stream);
int pmt_warp1d_mse_workspace(void);
int pmt_warp1d_mse_fwd_f32(const float* img, const float* off, const float* left, int mask_positive_disp,
                           double* workspace, float* loss, int N, int C, int H, int W, void* stream);
int pmt_warp1d_mse_bwd_f32(const float* img, const float* off, const float* left, int mask_positive_disp,
                           const float* gloss, float* gimg, float* goff, float* gleft, int N, int C, int H, int W,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * f4 (next row): SyncBatchNorm for a siamese pair fed as ONE batch x = [left; right] (2B,C,H,W).
 * The reference runs its tower twice per step (models/dsnet_t2.py:1159-1160) under nn.SyncBatchNorm
 * (torch_implementation.py:739): two invocations of every BN layer, each with its own batch statistics and its own
 * collective.  These kernels keep the two halves' statistics separate (same semantics, running statistics updated for
 * left then right) but let them travel in one all-gather / all-reduce per layer, which the caller (torch.distributed)
 * performs between the two calls of each direction:
 *   stats      : payload[(half*C+c)*2 + {0,1}] = mean, M2 = sum (x-mean)^2 of channel c in that half (local rank);
 *                payload[4C] = B*HW (this rank's element count); payload has 4C+1 floats
 *   apply      : gathered = [world][4C+1] (every rank's payload followed by its element count B*HW); combines them,
 *                out = (x-mean)*invstd*weight + bias, saves mean[2][C], invstd[2][C] (+ total count at invstd[2C]),
 *                updates running_mean/var (may be NULL) with momentum and the unbiased variance
 *   bwd_reduce : sums[(half*C+c)*2 + {0,1}] = sum dy, sum dy*(x-mean); gw/gb (C, ZEROED by the caller) += local grads
 *   bwd_apply  : sums after the cross-rank all-reduce; dx = (dy - sum_dy/N - (x-mean)*invstd^2*sum_dy_xmu/N)*invstd*weight
 * HW = H*W.  weight/bias may be NULL (non-affine).  `relu` is a flag word: bit 0 fuses the ReLU that follows every BN of
 * the reference's towers (norm -> relu -> conv): apply clamps at 0, the backward kernels gate dy where the recomputed y
 * was <= 0; bit 1 ("merged") treats the two halves as ONE batch -- plain SyncBatchNorm statistics over the whole batch,
 * one running-statistics update -- so the non-siamese layers of a model can use the same kernels and exchange.
 * ------------------------------------------------------------------------------------------- */
int pmt_bn_pair_stats_f32(const float* x, float* payload, int B, int C, int HW, void* stream);
int pmt_bn_pair_apply_f32(const float* x, const float* gathered, int world, const float* weight, const float* bias,
                          float* running_mean, float* running_var, float momentum, float eps, float* out,
                          float* save_mean, float* save_invstd, int B, int C, int HW, int relu, void* stream);
int pmt_bn_pair_bwd_reduce_f32(const float* dy, const float* x, const float* save_mean, const float* save_invstd,
                               float* sums, float* gw, float* gb, int B, int C, int HW, const float* weight,
                               const float* bias, int relu, void* stream);
int pmt_bn_pair_bwd_apply_f32(const float* dy, const float* x, const float* save_mean, const float* save_invstd,
                              const float* weight, const float* sums, float* dx, int B, int C, int HW,
                              const float* bias, int relu, void* stream);

/* The same four steps with the cross-rank exchange done BY THE KERNELS over NVLink peer memory instead of a collective
 * between them (north_star: NCCL only for the gradient all-reduce; the statistics of a layer are 4C+1 floats and a
 * collective launch per layer and direction is what limited the 8-GPU step).  Every rank owns a symmetric buffer of the
 * same layout, mapped into every peer (torch.distributed._symmetric_memory / cudaIpc); `peer_bufs` is a DEVICE array of
 * the `world` base pointers, `local_buf` this rank's own.  Per layer and direction the caller reserves, at float
 * offsets that are equal on all ranks, a payload region [2][world][n] (n = 4C+1 forward, 4C backward) and a flag
 * region [2][world] (int32, zero-initialised), plus LOCAL device words `epoch` (int, zero-initialised), `done`
 * (unsigned, zero-initialised) and a shared `err` word (set to 1 if a wait exceeded ~2 s).
 *   *_stats_peer / *_bwd_reduce_peer: compute into this rank's slot [epoch parity][rank]; the last block copies the slot
 *       into every peer's buffer, one thread fences at system scope, publishes the new epoch to every rank's flag
 *       (st.release.sys) and -- wait_peers != 0 -- waits until all `world` flags of local_buf show that epoch
 *       (wait_peers = 0 only when the ranks are emulated one after the other on ONE device: they must not wait for each
 *       other inside a kernel);
 *   *_apply_peer / *_bwd_apply_peer: read the payloads of all ranks from local_buf, no wait and no system-scope
 *       operation (backward: summed in rank order -> bit-identical on all ranks).
 * Epochs advance on the device, so the sequence can be captured in a CUDA graph and replayed. */
int pmt_bn_pair_stats_peer_f32(const float* x, void* const* peer_bufs, void* local_buf, int world, int rank,
                               int64_t payload_off, int64_t flag_off, int* epoch, unsigned* done, int* err, int wait_peers,
                               int B, int C, int HW, void* stream);
int pmt_bn_pair_apply_peer_f32(const float* x, void* local_buf, int world, int64_t payload_off, int64_t flag_off,
                               int* epoch, int* err, const float* weight, const float* bias, float* running_mean,
                               float* running_var, float momentum, float eps, float* out, float* save_mean,
                               float* save_invstd, int B, int C, int HW, int relu, void* stream);
int pmt_bn_pair_bwd_reduce_peer_f32(const float* dy, const float* x, const float* save_mean, const float* save_invstd,
                                    void* const* peer_bufs, void* local_buf, int world, int rank, int64_t payload_off,
                                    int64_t flag_off, int* epoch, unsigned* done, int* err, int wait_peers, float* gw,
                                    float* gb, int B, int C, int HW, const float* weight, const float* bias, int relu,
                                    void* stream);
int pmt_bn_pair_bwd_apply_peer_f32(const float* dy, const float* x, const float* save_mean, const float* save_invstd,
                                   const float* weight, void* local_buf, int world, int64_t payload_off,
                                   int64_t flag_off, int* epoch, int* err, float* dx, int B, int C, int HW,
                                   const float* bias, int relu, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer entry point for the headline workload (what a non-PyTorch caller of the reference's
 * sampler backend would bind): forward + backward of the 1 x P correlation on HOST tensors.
 * Copies in (in1,in2,gout), runs both kernels, copies out (out,gin1,gin2), in row blocks of ~1/8
 * batch item (image rows are independent) through 3 device slots with separate H2D / compute / D2H
 * streams (both PCIe directions and the kernels overlap), then synchronises.  Host buffers should
 * be pinned for full PCIe rate.
 * ------------------------------------------------------------------------------------------- */
int pmt_corr1d_fwd_bwd_host_f32(const float* in1_host, const float* in2_host, const float* gout_host,
                                float* out_host, float* gin1_host, float* gin2_host, int B, int C,
                                int H, int W, int P, int dilp);

/* ---------------------------------------------------------------------------------------------
 * Measurement helpers used by bench.py for the roofline denominators that MEASURED_PEAKS.json does
 * not carry.  pmt_probe_fp32_fma: runs a register-resident FFMA loop on every SM and returns the
 * achieved TFLOP/s in *tflops (timed with CUDA events on `stream`).  pmt_probe_copy: device copy
 * of `bytes` bytes src->dst with a float4 grid-stride kernel, returns GB/s (read+write).
 * ------------------------------------------------------------------------------------------- */
int pmt_probe_fp32_fma(int iters, double* tflops, void* stream);
int pmt_probe_copy(const void* src, void* dst, int64_t bytes, double* gbps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMT_OPS_H_ */
