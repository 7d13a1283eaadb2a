/* b200ode -- C ABI of the B200-native antisymmetric-ResNet Euler-block hot path.
 *
 * Drop-in boundary.  Every entry point replaces a piece of the
 * reference's TensorFlow graph for ONE path; citations are relative to the reference
 * repository (pierluigiferrari/differential_equations_resnet):
 *
 *   layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:85-155,210-293   kernel assembly (build)
 *   layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:157-171          call(): conv2d SAME + bias
 *   layers/tfkeras_layer_Conv2DAntisymmetric.py:90-175,216-270       general-k twin
 *   models/tfkeras_resnets.py:69-92                                  Euler step (conv, BN?, relu, h*, +x)
 *   training/training.py:300-301                                     backward (TF autodiff) + Adam
 *
 * Conventions
 *   - plain C, no C++/torch types; all tensor arguments are raw DEVICE pointers, NHWC,
 *     contiguous; shapes are explicit ints; `stream` is a cudaStream_t passed as void*.
 *   - every function returns 0 on success or a negative b200ode_status; the message is
 *     available through b200ode_last_error() (thread local).  Nothing aborts or throws.
 *   - stream-ordered and asynchronous: no hidden device synchronisation, no cudaMalloc / cudaFree
 *     inside a compute call.
 *   - the caller owns every tensor, the WORKSPACE included; the library owns only the opaque
 *     handles (staged weights, TMA descriptors).  Calls that need device scratch (the split-K
 *     partials of a weight gradient, the CUDA-core pre-activation buffer, the partials of the
 *     stem / transition / head gradients) take it from the block the caller bound to the handle
 *     (b200ode_layer_set_workspace / b200ode_chain_set_workspace) or passed with the call (glue
 *     entry points), sized by the *_workspace_bytes queries; a bound block that is too small is
 *     an error.  A handle with a bound block serves ONE call at a time.  With no block (NULL)
 *     the scratch comes from the device's stream-ordered memory pool (cudaMallocAsync /
 *     cudaFreeAsync on the call's stream: no synchronisation, graph capturable) and every call
 *     gets its own block, so calls on different streams -- also on one handle, after pack --
 *     never share scratch.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef B200ODE_H_
#define B200ODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ODE_VERSION 200

typedef enum {
  B200ODE_OK = 0,
  B200ODE_ERR_INVALID = -1,     /* bad argument / shape / alignment            */
  B200ODE_ERR_UNSUPPORTED = -2, /* configuration has no GPU kernel (no fallback) */
  B200ODE_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed          */
  B200ODE_ERR_NOT_PACKED = -4   /* compute call before b200ode_pack_kernel      */
} b200ode_status;

/* precision_mode */
#define B200ODE_PREC_STRICT 0     /* tcgen05 3xTF32 split, fp32 accumulate, fp32 I/O  (<=1e-5 rel) */
#define B200ODE_PREC_FAST_TF32 1  /* tcgen05 1xTF32, fp32 accumulate, fp32 I/O                     */
#define B200ODE_PREC_FAST_BF16 2  /* tcgen05 bf16 operands, fp32 accumulate, bf16 I/O              */
#define B200ODE_PREC_SIMT_FP32 3  /* CUDA-core fp32 FMA kernels (any C, any k, any stride)         */
#define B200ODE_PREC_FAST_F16  4  /* chains only: fp16 operands (11-bit significand = tf32 grade), rounded to nearest where they
                                     are produced, fp32 accumulate, fp32 residual stream; saved activations / dZ in fp16 */

/* param_layout: order of the free parameters in the flat packed vector = the
 * reference's variable creation order, each variable flattened C-order. */
#define B200ODE_LAYOUT_3BY3 0     /* [a,b,c,d (C each), W_0..W_{C-2} ([3,3,C-o-1]), bias]  (3By3.py:119-153,219-245) */
#define B200ODE_LAYOUT_GENERAL 1  /* per o: diag scalars, W_o [k,k,C-o-1,1]; bias last (Conv2DAntisymmetric.py:117-157) */

/* fuse_flags of the Euler forward/backward */
#define B200ODE_F_BIAS 1       /* z = conv + bias                                 */
#define B200ODE_F_RELU 2       /* r = relu(z)                                     */
#define B200ODE_F_SCALE 4      /* r = h * r      (models/tfkeras_resnets.py:90-91) */
#define B200ODE_F_RESIDUAL 8   /* y = r + x      (models/tfkeras_resnets.py:92)    */
#define B200ODE_F_EULER (1 | 2 | 4 | 8)

typedef struct b200ode_layer b200ode_layer_t;

int b200ode_version(void);
const char* b200ode_last_error(void);
/* 1 when a CUDA device of compute capability 10.x is usable, else 0 (message set). */
int b200ode_device_ok(void);

/* ---- layer handle: replaces Conv2DAntisymmetric3By3.__init__/build (3By3.py:60-155) and
 *      Conv2DAntisymmetric.__init__/build (Conv2DAntisymmetric.py:60-159) ---- */
int b200ode_layer_create(int channels, int ksize, float gamma, int stride_h, int stride_w, int use_bias,
                         int antisymmetric, int precision_mode, int param_layout, b200ode_layer_t** out);
int b200ode_layer_destroy(b200ode_layer_t* layer);
/* number of floats in the packed parameter vector (bias included when use_bias) */
int64_t b200ode_layer_num_params(const b200ode_layer_t* layer);
/* precision mode actually in effect (SIMT is selected for shapes the tensor path cannot take) */
int b200ode_layer_effective_mode(const b200ode_layer_t* layer);

/* Workspace of the layer's compute calls for inputs of shape [N,H,W,C] (max over forward, data and weight gradient;
 * 0 when none is needed) and binding of a caller-owned, 256-byte aligned device block (NULL unbinds).  SURVEY.md 8b. */
int b200ode_layer_workspace_bytes(const b200ode_layer_t* layer, int N, int H, int W, size_t* bytes_out);
int b200ode_layer_set_workspace(b200ode_layer_t* layer, void* workspace, size_t bytes);

/* ---- K1 antisym_pack: replaces the assembly graph 3By3.py:113-141,268-293 (re-executed every
 *      sess.run by the reference).  params: device fp32 [num_params].  Writes the handle's
 *      staged tensor-core operand copies and, when K_dense_hwio != NULL, the dense fp32
 *      [k,k,C,C] kernel that get_kernel() (3By3.py:188-199) returns. ---- */
int b200ode_pack_kernel(b200ode_layer_t* layer, const float* params, float* K_dense_hwio, void* stream);

/* ---- K2 forward: tf.nn.conv2d(SAME)+bias (3By3.py:159-169) fused with relu, h*, +x
 *      (models/tfkeras_resnets.py:89-92) according to fuse_flags.
 *      x, y: [N,H,W,C] (fp32, or bf16 in FAST_BF16 mode); y: [N,ceil(H/sh),ceil(W/sw),C].
 *      relu_mask: nullable; when given receives 1 bit per output element, row pitch
 *      ceil(C/8) bytes per pixel, bit c%8 of byte c/8 = (z > 0).
 *      z_out: nullable fp32 pre-activation output (needed by the BatchNorm path). ---- */
int b200ode_euler_fwd(b200ode_layer_t* layer, const void* x, void* y, uint8_t* relu_mask, float* z_out, int N, int H,
                      int W, float h, int fuse_flags, void* stream);

/* ---- K3 data gradient (TF autodiff Conv2DBackpropInput, training/training.py:300):
 *      dx = [dy_skip] + conv_transpose_K(dz).  For the antisymmetric kernel this is
 *      dy_skip - conv_K(dz) + 2*gamma*dz, i.e. the forward kernel with the forward weights.
 *      dz: [N,Ho,Wo,C]; dy_skip: nullable [N,H,W,C] added to the result; dx: [N,H,W,C]. ---- */
int b200ode_euler_dgrad(b200ode_layer_t* layer, const void* dz, const void* dy_skip, void* dx, int N, int H, int W,
                        void* stream);
/* The same data gradient fused with the relu/h backward of the step BELOW it in a chain of Euler steps (the TF autodiff
 * of models/tfkeras_resnets.py:89-92 for the previous block): in one pass
 *   dx      = dy_skip - conv_K(dz) + 2*gamma*dz
 *   dz_prev = h * dx * prev_relu_mask          (bit-identical to b200ode_euler_dgrad followed by b200ode_relu_scale_bwd) */
int b200ode_euler_dgrad_fused(b200ode_layer_t* layer, const void* dz, const void* dy_skip, void* dx,
                              const uint8_t* prev_relu_mask, void* dz_prev, float h, int N, int H, int W, void* stream);

/* ---- K4 weight gradient (Conv2DBackpropFilter + backprop through the assembly graph):
 *      dense G = sum_pixels x^T dz, folded onto the free parameters
 *      (S = G - rot180(G)^T); grad_params: fp32 [num_params] (bias
 *      gradient = sum_pixels dz in the bias slot).  G_dense: nullable fp32 [k,k,C,C].
 *      accumulate != 0 adds into grad_params instead of overwriting. ---- */
int b200ode_euler_wgrad(b200ode_layer_t* layer, const void* x, const void* dz, float* grad_params, float* G_dense,
                        int N, int H, int W, int accumulate, void* stream);

/* ---- elementwise / reduction tails (K5/K6) ---- */
/* dz = h * dy * mask   (backward of relu + Lambda(h*x)); dtype = layer I/O dtype */
int b200ode_relu_scale_bwd(const void* dy, const uint8_t* relu_mask, void* dz, int64_t pixels, int channels, float h,
                           int is_bf16, void* stream);
/* y = h * relu(z * scale[c] + shift[c]) + x ; optional mask (same format as euler_fwd) */
int b200ode_euler_tail(const float* z, const float* scale, const float* shift, const float* x, float* y,
                       uint8_t* relu_mask, int64_t pixels, int channels, float h, int fuse_flags, void* stream);
/* per-channel sums over pixels: out_sum[c] = sum a[p,c], out_sumsq[c] = sum a[p,c]*b[p,c] (b nullable -> a*a).
 * workspace: device fp32 [2 * B200ODE_COLSUM_PARTS * channels]. Deterministic. */
#define B200ODE_COLSUM_PARTS 2048
int b200ode_colsum(const float* a, const float* b, float* out_sum, float* out_sumprod, float* workspace, int64_t pixels,
                   int channels, void* stream);
/* BatchNorm Euler step (models/tfkeras_resnets.py:70-87), forward part 1: z_out = conv_K(x) + b as fp32 [N,H,W,C] and, from
 * the epilogue registers of the SAME kernel (no second pass over z), per-channel partial sums of z and z*z:
 * stats_ws = device fp32 [2 * B200ODE_COLSUM_PARTS * C] receives *rows_out (host int) rows of sums, then as many rows of
 * sums of squares.  fp32 modes only (STRICT / FAST_TF32 / SIMT). */
int b200ode_euler_fwd_bn_stats(b200ode_layer_t* layer, const void* x, float* z_out, float* stats_ws, int* rows_out, int N, int H,
                               int W, void* stream);
/* part 2: fixed-order reduction of the rows (deterministic) into out_sum / out_sumsq (nullable, [C] each) and, when
 * bn_gamma != NULL, b200ode_bn_finalize over `pixels` in the same launch.  Data parallel SyncBN: call with bn_gamma = NULL,
 * all-reduce the 2C sums, then b200ode_bn_finalize with the global pixel count. */
int b200ode_bn_stats_finalize(const float* stats_ws, int rows, float* out_sum, float* out_sumsq, const float* bn_gamma,
                              const float* bn_beta, float* mean, float* inv_std, float* scale, float* shift, float* moving_mean,
                              float* moving_var, int64_t pixels, int channels, float eps, float momentum, void* stream);
/* training-mode BatchNormalization(axis=3) (models/tfkeras_resnets.py:85-87; Keras eps 1e-3):
 * from sums -> mean, inv_std, and the affine (scale, shift) that b200ode_euler_tail consumes;
 * updates moving statistics (momentum) when moving_mean != NULL. */
int b200ode_bn_finalize(const float* sum, const float* sumsq, const float* bn_gamma, const float* bn_beta, float* mean,
                        float* inv_std, float* scale, float* shift, float* moving_mean, float* moving_var,
                        int64_t pixels, int channels, float eps, float momentum, void* stream);
/* BN backward, two passes: (1) du = h*dy*[u>0] with u = z*scale+shift; sums of du and du*zhat
 * via b200ode_colsum-like reduction into dbeta/dgamma; (2) dz. */
int b200ode_bn_bwd_reduce(const float* dy, const float* z, const float* scale, const float* shift, const float* mean,
                          const float* inv_std, float* dgamma, float* dbeta, float* workspace, int64_t pixels,
                          int channels, float h, void* stream);
int b200ode_bn_bwd_apply(const float* dy, const float* z, const float* scale, const float* shift, const float* mean,
                         const float* inv_std, const float* bn_gamma, const float* dgamma, const float* dbeta,
                         float* dz, int64_t pixels, int channels, float h, void* stream);

/* ---- optimiser step: tf.train.AdamOptimizer (training/training.py:300-301; eps 1e-7 in the
 *      notebooks).  step_counter: device int32, 1-based step of THIS update (read, not written).
 *      grad_scale multiplies the gradient first (1/world_size after an all-reduce sum). ---- */
int b200ode_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, const int32_t* step_counter,
                      float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
int b200ode_increment(int32_t* counter, void* stream);
/* The reference trainer's gradient metric (training/training.py:385-407, `<layer>_kernel_gradient_mean_norm`):
 * out[i] = || grad_scale * grads[offsets[i] .. offsets[i]+sizes[i]) ||_2 / sizes[i], one launch for all layers
 * (offsets / sizes are DEVICE arrays of n_slices int64). */
int b200ode_gradient_mean_norms(const float* grads, const int64_t* offsets, const int64_t* sizes, int n_slices,
                                float grad_scale, float* out, void* stream);

/* ---- persistent Euler-step chains: the stage loop of models/tfkeras_resnets.py:575-593
 *      (n x single_layer_identity_block, :28-94, on a tensor of constant shape) forward and its
 *      TF-autodiff backward sweep (training/training.py:300) in ONE launch per direction.  One CTA
 *      keeps one image in shared memory across all steps.  Shapes whose image does not fit shared
 *      memory are refused (b200ode_chain_supported == 0): use the per-layer calls.
 *      Three formulations, chosen by precision_mode at create time:
 *        STRICT     3xTF32 inside the chain kernel (fp32-grade: <= 1e-5 relative of the reference's fp32 arithmetic): the fp32
 *                   residual strip is the hi operand, the epilogue also writes its tf32 remainder strip, every tap carries
 *                   W_hi and W_lo rows; saved activations and dZ are fp32 (same buffers as FAST_TF32), weight gradient by
 *                   the strict layer-batched kernel.
 *        FAST_TF32  tf32 operands straight from the fp32 residual stream (8 channels per MMA); saved
 *                   activations and dZ are fp32; forward results equal the per-layer FAST_TF32 kernels bit for bit.
 *        FAST_F16   fp16 operands (16 channels per MMA, same 11-bit significand) rounded to nearest where they are
 *                   produced, fp32 residual stream kept in shared memory; saved operands (acts, dz_all) are fp16
 *                   and the backward strips carry ONE power-of-two scale per launch derived from max|dy|
 *                   (fp16 range); b200ode_chain_wgrad undoes it.  |activations| must stay below 65504. ---- */
typedef struct b200ode_chain b200ode_chain_t;
int b200ode_chain_supported(int channels, int H, int W, int precision_mode);
/* n_layers distinct antisymmetric 3x3 layers (LAYOUT_3BY3 parameters, strides (1,1)) */
int b200ode_chain_create(int channels, int n_layers, float gamma, int use_bias, int precision_mode,
                         b200ode_chain_t** out);
int b200ode_chain_destroy(b200ode_chain_t* chain);
int64_t b200ode_chain_layer_params(const b200ode_chain_t* chain);
/* workspace of b200ode_chain_wgrad (split-K partials of all layers) for [N,H,W,C] images; binding as for layers */
int b200ode_chain_workspace_bytes(const b200ode_chain_t* chain, int N, int H, int W, size_t* bytes_out);
int b200ode_chain_set_workspace(b200ode_chain_t* chain, void* workspace, size_t bytes);
/* K1 for all layers at once: params + l*param_layer_stride = packed parameters of layer l */
int b200ode_chain_pack(b200ode_chain_t* chain, const float* params, int64_t param_layer_stride, void* stream);
/* n_steps Euler steps x_{l+1} = x_l + h*relu(conv_{K_l}(x_l)+b_l); step l uses layer l % n_layers
 * (n_layers == 1: the long-horizon integration through one block).  relu_masks: nullable
 * [n_steps][N,H,W,C/8]; y_final: [N,H,W,C] fp32 output of the last step.
 *   FAST_TF32: acts nullable fp32 [n_steps][N,H,W,C] receiving every x_{l+1}; y_final nullable (required when acts is NULL).
 *   FAST_F16:  acts nullable fp16 [n_steps][N,H,W,C] receiving the INPUT of every step (acts[0] = fp16(x0)): the
 *              weight-gradient operands; y_final required. */
int b200ode_chain_fwd(b200ode_chain_t* chain, const float* x0, void* acts, uint8_t* relu_masks, float* y_final, int N,
                      int H, int W, float h, int n_steps, void* stream);
/* backward sweep over the n_layers steps: dy = dL/dx_L -> dx = dL/dx_0 (both fp32); dz_all
 * [n_layers][N,H,W,C] receives dZ_l = h*dY_l*mask_l, the input of the weight gradient: fp32 (FAST_TF32) or
 * fp16 times the launch's power-of-two scale (FAST_F16; the scale lives in the chain handle). */
int b200ode_chain_dgrad(b200ode_chain_t* chain, const float* dy, const uint8_t* relu_masks, void* dz_all, float* dx,
                        int N, int H, int W, float h, void* stream);
/* Same with max|dy| supplied by the caller (device scalar, written by the kernel that produced dy: the *_amax entries of the
 * head / transition data gradients below): FAST_F16 chains skip their own reduction over dy (two launches less on the
 * critical path of a train step); dy_amax must stay untouched until b200ode_chain_wgrad of this chain has run.  Other modes
 * ignore it. */
int b200ode_chain_dgrad_amax(b200ode_chain_t* chain, const float* dy, const uint8_t* relu_masks, void* dz_all, float* dx,
                             int N, int H, int W, float h, const float* dy_amax, void* stream);
/* weight + bias gradients of all layers in one launch (+ fold/reduce launches); result at
 * grad_params + l*grad_layer_stride.
 *   FAST_TF32: layer l reads x_l (x0 for l = 0, acts[l-1] otherwise) and dz_all[l], all fp32.
 *   FAST_F16:  layer l reads acts[l] and dz_all[l] (fp16, as written by the two calls above; x0 is ignored). */
int b200ode_chain_wgrad(b200ode_chain_t* chain, const float* x0, const void* acts, const void* dz_all,
                        float* grad_params, int64_t grad_layer_stride, int N, int H, int W, void* stream);

/* ---- stem / transition / head of the single-block ResNet, fp32 CUDA-core kernels.
 *      All tensors NHWC fp32 (images: uint8 or fp32); kernels in the Keras HWIO layout.
 *      The three gradient entry points take their scratch as (workspace, workspace_bytes): a device block of at least
 *      b200ode_glue_workspace_bytes(op, ...) bytes, or NULL for the stream-ordered pool. ---- */
#define B200ODE_GLUE_STEM_WGRAD 0        /* (N,H,W,Cin,Cout) of b200ode_stem_wgrad; strides ignored            */
#define B200ODE_GLUE_TRANSITION_WGRAD 1  /* (N,H,W,Cin,Cout,stride_h,stride_w) of b200ode_transition_wgrad      */
#define B200ODE_GLUE_HEAD 2              /* N, Cin = channels, Cout = classes of b200ode_head_fwd_bwd; H,W >= 1 */
int b200ode_glue_workspace_bytes(int op, int N, int H, int W, int Cin, int Cout, int stride_h, int stride_w,
                                 size_t* bytes_out);
/* input Lambda layers + conv1 + relu (models/tfkeras_resnets.py:555-572), 3x3, strides (1,1):
 * out = relu(conv_SAME((images - subtract_mean) / divide_by_stddev) + bias)   [normalize == 0: raw images] */
int b200ode_stem_fwd(const void* images, int images_are_u8, float subtract_mean, float divide_by_stddev, int normalize,
                     const float* kernel_hwio, const float* bias, float* out, int N, int H, int W, int Cin, int Cout,
                     void* stream);
/* dparams = [dkernel (3,3,Cin,Cout) | dbias (Cout)] from dout = dL/dout and the saved stem output */
int b200ode_stem_wgrad(const void* images, int images_are_u8, float subtract_mean, float divide_by_stddev, int normalize,
                       const float* out, const float* dout, float* dparams, int N, int H, int W, int Cin, int Cout,
                       void* workspace, size_t workspace_bytes, void* stream);
/* single_layer_conv_block (models/tfkeras_resnets.py:204-269): out = relu(conv3x3_s(x)+bm) + conv1x1_s(x)+bs,
 * TF SAME padding (pad_before = total/2); relu_mask: 1 bit per output element ((main > 0), as euler_fwd).
 * Strides (2,2) with 16 -> 32 or 32 -> 64 channels (the reference's transitions) run as tensor-core implicit GEMMs with the
 * 3xTF32 split (fp32-grade, <= 1e-5; csrc/kernels_glue_mma.cuh), every other shape on fp32 CUDA-core kernels. */
int b200ode_transition_fwd(const float* x, const float* main_kernel, const float* main_bias, const float* short_kernel,
                           const float* short_bias, float* out, uint8_t* relu_mask, int N, int H, int W, int Cin, int Cout,
                           int stride_h, int stride_w, void* stream);
int b200ode_transition_dgrad(const float* dout, const uint8_t* relu_mask, const float* main_kernel,
                             const float* short_kernel, float* dx, int N, int H, int W, int Cin, int Cout, int stride_h,
                             int stride_w, void* stream);
/* ... and *dx_amax = max(*dx_amax, max|dx|) as a side effect (the caller zeroes the scalar beforehand; fed to
 * b200ode_chain_dgrad_amax of the stage in front of the transition) */
int b200ode_transition_dgrad_amax(const float* dout, const uint8_t* relu_mask, const float* main_kernel,
                                  const float* short_kernel, float* dx, int N, int H, int W, int Cin, int Cout, int stride_h,
                                  int stride_w, float* dx_amax, void* stream);
/* dparams = [dmain_kernel (3,3,Cin,Cout) | dmain_bias | dshort_kernel (Cin,Cout) | dshort_bias] */
int b200ode_transition_wgrad(const float* x, const float* dout, const uint8_t* relu_mask, float* dparams, int N, int H, int W,
                             int Cin, int Cout, int stride_h, int stride_w, void* workspace, size_t workspace_bytes,
                             void* stream);
/* The same three calls for the FAST precision modes (reference models/tfkeras_resnets.py:204-269 under a stated tolerance
 * instead of 1e-5): on the tensor-core shapes (stride 2, 16 -> 32 and 32 -> 64 channels) ONE tf32 MMA on operands rounded to
 * nearest -- the grade of the fast modes' fp16 / tf32 / bf16 chain operands -- replaces the 3xTF32 split (measured at cfg3:
 * fwd 16.9 -> 12.6 us, dgrad 17.7 -> 10.7 us, wgrad 29.1 -> 22.7 us); every other shape runs the fp32 kernels of the plain
 * entries.  b200ode_transition_dgrad_fast takes the optional dx_amax of b200ode_transition_dgrad_amax (NULL: none). */
int b200ode_transition_fwd_fast(const float* x, const float* main_kernel, const float* main_bias, const float* short_kernel,
                                const float* short_bias, float* out, uint8_t* relu_mask, int N, int H, int W, int Cin, int Cout,
                                int stride_h, int stride_w, void* stream);
int b200ode_transition_dgrad_fast(const float* dout, const uint8_t* relu_mask, const float* main_kernel,
                                  const float* short_kernel, float* dx, int N, int H, int W, int Cin, int Cout, int stride_h,
                                  int stride_w, float* dx_amax, void* stream);
int b200ode_transition_wgrad_fast(const float* x, const float* dout, const uint8_t* relu_mask, float* dparams, int N, int H, int W,
                                  int Cin, int Cout, int stride_h, int stride_w, void* workspace, size_t workspace_bytes,
                                  void* stream);
/* GlobalAveragePooling2D -> Dense(softmax) (models/tfkeras_resnets.py:595-597) -> mean
 * K.categorical_crossentropy(onehot, probs) with clipping eps (training/training.py:295), forward and
 * backward in one call: loss (1 float), dx = dL/dx [N,HW,C] (nullable), dparams = [dfc_kernel (C,K) | dfc_bias]
 * (nullable), probs [N,K] (nullable). */
int b200ode_head_fwd_bwd(const float* x, const float* fc_kernel, const float* fc_bias, const float* onehot, float eps,
                         float* probs, float* loss, float* dx, float* dparams, int N, int HW, int C, int K, void* workspace,
                         size_t workspace_bytes, void* stream);
/* ... and *dx_amax = max(*dx_amax, max|dx|) as a side effect (dx required; see b200ode_chain_dgrad_amax) */
int b200ode_head_fwd_bwd_amax(const float* x, const float* fc_kernel, const float* fc_bias, const float* onehot, float eps,
                              float* probs, float* loss, float* dx, float* dparams, int N, int HW, int C, int K, void* workspace,
                              size_t workspace_bytes, float* dx_amax, void* stream);

/* ---- gradient exchange ------------------------------------------------------------
 * The reference has no collective (single tf.Session, training/training.py:132); data parallel training sums
 * the flat packed-gradient bucket over ranks.  NCCL is bound at run time (dlopen of libnccl.so.2, the copy the
 * process already holds if a framework loaded one), so the library itself has no link-time NCCL dependency.
 * One process per GPU; the caller distributes the 128-byte unique id (rank 0 creates it) out of band. */
typedef struct b200ode_comm b200ode_comm_t;
#define B200ODE_UNIQUE_ID_BYTES 128
int b200ode_comm_unique_id(void* id_out /* B200ODE_UNIQUE_ID_BYTES bytes */);
int b200ode_comm_init(int nranks, int rank, const void* nccl_unique_id, b200ode_comm_t** out);
/* in-place sum of buf[0..n) (device pointer, fp32) over all ranks, stream ordered, graph capturable */
int b200ode_comm_allreduce_bucket(b200ode_comm_t* comm, float* buf, size_t n, void* stream);
/* Peer-memory exchange fused into the optimiser (one box, NVLink / NVSwitch, up to B200ODE_MAX_RANKS ranks):
 * b200ode_comm_shared_alloc (collective, once per communicator) allocates this rank's gradient bucket of n_floats inside
 * the library, maps it into every other rank (CUDA IPC; the handles travel through the communicator) and returns the
 * LOCAL pointer: the caller's weight-gradient kernels write their gradients there.  b200ode_comm_adam_step is
 * b200ode_adam_step over a slice of that bucket (grads_local points into it; offset and n multiples of 4) whose
 * gradient is the SUM over ranks, read straight from peer memory inside the Adam kernel in rank order (bit-identical
 * parameters on every rank) and scaled by 1/nranks: no separate all-reduce launch.  Stream ordered, graph capturable;
 * every rank must issue the same sequence of calls.  The kernel returns only after all ranks have read this rank's
 * slice, so the next step may overwrite it.
 * Two-shot form: the region also holds a parameter replica (*params_out, nullable).  When `params` of
 * b200ode_comm_adam_step points into it at the gradient's offset, every rank sums and updates only its 1/nranks shard
 * of the slice (its m, v entries are the only ones touched: sharded optimiser state) and stores the new parameters
 * into every rank's replica: (nranks-1)/nranks of the slice crosses NVLink each way instead of (nranks-1) x. */
#define B200ODE_MAX_RANKS 8
int b200ode_comm_shared_alloc(b200ode_comm_t* comm, size_t n_floats, float** local_out, float** params_out);
int b200ode_comm_adam_step(b200ode_comm_t* comm, float* params, const float* grads_local, float* m, float* v, int64_t n,
                           float lr, float beta1, float beta2, float eps, const int32_t* step_counter, void* stream);
int b200ode_comm_destroy(b200ode_comm_t* comm);

/* test hook: number of kernel launches issued by this library in this process */
int64_t b200ode_launch_count(void);
/* debug hook, host only (no device needed): the tile plan the per-layer convolution would use;
 * out8 = {images per tile, 128-position segments per image per tile, tiles per image, total tiles, grid, A stages,
 * weight stages, accumulator stages (2 = the drain of a tile overlaps the next tile's MMAs)}. */
int b200ode_debug_conv_plan(int precision_mode, int channels, int N, int H, int W, int* out8);
/* debug hook: device buffer of uint64 [ctas][16] that the tensor-core kernels fill with a per-CTA
 * timeline (slot 0/15: %globaltimer ns at CTA start/end; others: SM clock deltas); NULL disables. */
int b200ode_debug_set_trace(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* B200ODE_H_ */
