/*
 * l2i_b200.h - C ABI of the B200-native (sm_100a) Latent2im hot path.
 *
 * The drop-in boundary of this repository.  Every entry point takes plain pointers and sizes
 * (device pointers unless a parameter says "host"), enqueues work on the CUDA stream it is given
 * (a cudaStream_t passed as void*), never synchronises, never allocates inside a compute call and
 * never throws: it returns L2I_OK (0) or a negative L2I_ERR_* code, and l2i_last_error_string()
 * describes the most recent failure on the calling thread.
 *
 * Each group cites the reference interface (KelestZ/Latent2im, paths under
 * graphs/stylegan_v2_real/) it replaces.  INTEGRATION.md shows the binding a maintainer of the
 * reference would add on top of this header.
 */
#ifndef L2I_B200_H
#define L2I_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L2I_OK 0
#define L2I_ERR_INVALID_ARG (-1)   /* bad size / null pointer / unsupported combination */
#define L2I_ERR_CUDA (-2)          /* a CUDA runtime or driver call failed               */
#define L2I_ERR_UNSUPPORTED (-3)   /* valid request this build has no kernel for         */
#define L2I_ERR_STATE (-4)         /* generator used before finalize, batch > max_batch… */

/* arithmetic / storage types */
#define L2I_F32 0
#define L2I_BF16 1
#define L2I_F16 2

/* -------------------------------------------------------------------------------------------- */
/* library                                                                                       */
/* -------------------------------------------------------------------------------------------- */
int l2i_abi_version(void);
const char* l2i_last_error_string(void);
/* number of kernels this library has launched since load (bench.py's "gpu_launches") */
int64_t l2i_launch_count(void);

/* -------------------------------------------------------------------------------------------- */
/* native-op drop-ins                                                                            */
/* -------------------------------------------------------------------------------------------- */

/* Replaces fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 * (op/fused_bias_act.cpp:11-21, kernel op/fused_bias_act_kernel.cu:18-49).
 *   y[i] = act_grad(x[i] + bias[(i / step_b) % size_b], ref[i]) * scale
 * bias may be NULL (size_b == 0); ref may be NULL unless grad == 1.  act: 1 linear, 3 leaky relu.
 * n is a 64-bit element count (the reference's int32 size_x overflows at 2^31). */
int l2i_fused_bias_act(void* y, const void* x, const void* bias, const void* ref,
                       int64_t n, int64_t step_b, int64_t size_b,
                       int act, int grad, float alpha, float scale, int dtype, void* stream);

/* Fused backward of fused_leaky_relu (op/fused_act.py:19-39): grad_in = act'(out) * grad_out * scale
 * and grad_bias[c] = sum over batch and spatial of grad_in, accumulated in fp32.
 * x is viewed as [outer, size_b, step_b]. grad_bias (fp32, size_b) may be NULL. */
int l2i_fused_leaky_relu_bwd(void* grad_in, float* grad_bias, const void* grad_out, const void* out,
                             int64_t outer, int64_t size_b, int64_t step_b,
                             float alpha, float scale, int dtype, void* stream);

/* Replaces upfirdn2d_op.upfirdn2d(input[major,in_h,in_w,minor], kernel[kh,kw], up_x, up_y, down_x,
 * down_y, pad_x0, pad_x1, pad_y0, pad_y1) (op/upfirdn2d.cpp:12-22, op/upfirdn2d_kernel.cu:52-272).
 * kernel is fp32 on the device, kh,kw <= 8.  Any up/down/pad combination is computed (the
 * reference silently returns uninitialised memory outside its six "modes").
 * y has shape [major, out_h, out_w, minor] with
 *   out_h = (in_h*up_y + pad_y0 + pad_y1 - kh + down_y) / down_y   (same for w). */
int l2i_upfirdn2d(void* y, const void* x, const float* kernel,
                  int64_t major, int in_h, int in_w, int minor, int kh, int kw,
                  int up_x, int up_y, int down_x, int down_y,
                  int pad_x0, int pad_x1, int pad_y0, int pad_y1, int dtype, void* stream);

/* -------------------------------------------------------------------------------------------- */
/* latent side: mapping network, walk modules                                                    */
/* -------------------------------------------------------------------------------------------- */

/* Generic fused linear used by the mapping network, the MLP walks and their backward:
 *   y[b, n] = act( wscale * sum_k x[b, k] * W[n, k] + bscale * bias[n] ) * gain
 * x: [B, K] with row stride x_stride; W: [N, K] row-major; y: [B, N] with row stride y_stride.
 * act: 0 none, 1 leaky relu (slope alpha).  All fp32.
 * Replaces F.linear + fused_leaky_relu in EqualLinear.forward (networks.py:148-156) and the
 * nn.Linear + nn.LeakyReLU pairs of the walk MLPs (transform_base.py:175-179, 214-217). */
int l2i_linear_fwd(float* y, int64_t y_stride, const float* x, int64_t x_stride,
                   const float* W, const float* bias, int B, int N, int K,
                   float wscale, float bscale, int act, float alpha, float gain, void* stream);

/* PixelNorm (networks.py:11-16): y[b,:] = x[b,:] * rsqrt(mean(x[b,:]^2) + 1e-8). fp32. */
int l2i_pixel_norm(float* y, const float* x, int B, int D, void* stream);

/* WalkLinearMultiW.forward (transform_base.py:151-165) for all latent layers in one launch:
 *   out[b, i, :] = in[b, i, :] + (layer_mask >> i & 1 ? sum_a alpha[b, a] * w[a, i, :] : 0)
 * in/out: [B, n_latent, D] fp32 (in_layer_stride may be 0 when the W+ list is one tensor
 * repeated, transform_base.py:372-378); alpha: [B, A]; w: [A, n_latent, D]. */
int l2i_walk_linear_fwd(float* out, const float* in, int64_t in_batch_stride, int64_t in_layer_stride,
                        const float* alpha, const float* w, int B, int A, int n_latent, int D,
                        uint64_t layer_mask, void* stream);

/* Gradient of the above w.r.t. the walk parameter and the input latents:
 *   grad_w[a, i, :] = sum_b alpha[b, a] * grad_out[b, i, :]   (masked layers get 0)
 *   grad_in         = grad_out                                 (identity; not materialised) */
int l2i_walk_linear_bwd(float* grad_w, const float* grad_out, const float* alpha,
                        int B, int A, int n_latent, int D, uint64_t layer_mask, void* stream);

/* out[b, i, :] = in[b, i, :] + coef[b] * d[b, i, :]           (WalkMlpMultiW, :188-193)
 * or, with normalize = 1,  in + d / ||d||_2                   (WalkNonLinearW, :227-229)
 * d_layer_stride may be 0 when one MLP output serves every layer. */
int l2i_walk_combine(float* out, const float* in, int64_t in_batch_stride, int64_t in_layer_stride,
                     const float* d, int64_t d_batch_stride, int64_t d_layer_stride,
                     const float* coef, int B, int n_latent, int D, uint64_t layer_mask,
                     int normalize, void* stream);

/* Backward of l2i_linear_fwd (wscale = bscale = gain = 1) for training the walk MLPs, i.e. what autograd runs for
 * nn.Linear + nn.LeakyReLU in WalkMlpMultiW / WalkNonLinearW (transform_base.py:175-179, 214-217):
 *   g = gy * (act && y <= 0 ? alpha : 1)      (y = the saved forward OUTPUT)
 *   gx[b, k] = sum_n g[b, n] W[n, k];   gW[n, k] = sum_b g[b, n] x[b, k];   gb[n] = sum_b g[b, n]
 * gx / gW / gb may each be NULL (not wanted).  All fp32, dense row-major. */
int l2i_linear_bwd(float* gx, float* gW, float* gb, const float* gy, const float* y, const float* x,
                   const float* W, int B, int N, int K, int act, float alpha, void* stream);

/* Gradient of l2i_walk_combine w.r.t. d (grad w.r.t. `in` is grad_out itself):
 *   normalize = 0: grad_d = coef * g;    normalize = 1: grad_d = (g - <g,u> u) / ||d||, u = d / ||d||
 * grad_out: [B, n_latent, D].  d_layer_stride == 0 (one MLP output for every layer): grad_d is [B, D] = the sum over
 * the masked layers and `scratch` ([B, n_latent, D]) is required; otherwise grad_d is [B, n_latent, D]. */
int l2i_walk_combine_bwd(float* grad_d, float* scratch, const float* grad_out, const float* d,
                         int64_t d_batch_stride, int64_t d_layer_stride, const float* coef, int B,
                         int n_latent, int D, uint64_t layer_mask, int normalize, void* stream);

/* -------------------------------------------------------------------------------------------- */
/* StyleGAN2 synthesis network                                                                   */
/* -------------------------------------------------------------------------------------------- */

typedef struct l2i_generator l2i_generator_t;

/* Replaces Generator.__init__ (networks.py:361-438) on the native side.
 * dtype is the activation / tensor-core operand type: L2I_F32 (CUDA-core fp32 path, the parity
 * arbiter on the GPU) or L2I_BF16 (tcgen05 path, fp32 accumulation in TMEM).
 * Workspace for batches up to max_batch is allocated here, never during forward. */
int l2i_generator_create(l2i_generator_t** out, int size, int style_dim, int n_mlp,
                         int channel_multiplier, const float* blur_taps /* host */, int n_blur_taps,
                         float lr_mlp, int dtype, int max_batch);
void l2i_generator_destroy(l2i_generator_t* g);

/* Uploads one tensor of the rosinality-format state_dict by its key ("convs.3.conv.weight", ...)
 * from a DEVICE fp32 buffer of numel elements.  Unknown keys return L2I_ERR_INVALID_ARG. */
int l2i_generator_set_param(l2i_generator_t* g, const char* key, const float* data, int64_t numel,
                            void* stream);
/* Repacks weights into the kernel layouts (scaled, transposed, bf16 copies, sum-of-squares for
 * demodulation).  Must be called after the last set_param and before forward. */
int l2i_generator_finalize(l2i_generator_t* g, void* stream);

int l2i_generator_num_layers(const l2i_generator_t* g);
int l2i_generator_n_latent(const l2i_generator_t* g);

/* Replaces Generator.style (networks.py:374-382, 457-458): z[B, D] -> w[B, D], fp32. */
int l2i_generator_mapping(l2i_generator_t* g, float* w, const float* z, int batch, void* stream);

/* Replaces Generator.forward(styles=latent, input_is_latent=True, noise=[...])
 * (networks.py:460-514).
 *   latent : [B, n_latent, D] fp32, with explicit strides in elements
 *   noise  : HOST array of num_layers DEVICE pointers, fp32, each [Bn, 1, H_i, W_i] with
 *            noise_batch[i] = Bn in {1, B} (1 = broadcast over the batch, as the registered
 *            noise buffers are)
 *   image  : [B, 3, size, size] fp32 NCHW (the reference's return value), may be NULL
 *   image_u8: optional [B, size, size, 3] uint8 = clip((x+1)/2*255, 0, 255) truncated
 *            (transform_base.py:625-626 + the NCHW->NHWC transpose of :655), may be NULL
 */
int l2i_generator_forward(l2i_generator_t* g, const float* latent, int64_t latent_batch_stride,
                          int64_t latent_layer_stride, const float* const* noise,
                          const int* noise_batch, float* image, uint8_t* image_u8, int batch,
                          void* stream);

/* Training mode: while enabled, forward() additionally keeps every layer's activation (and the raw
 * up-conv outputs) in per-layer buffers so that backward() can run.  The first call allocates those
 * buffers and the transposed weight copies and synchronises the device; forward() still never
 * allocates.  The noise tensors passed to a training forward must stay alive until its backward. */
int l2i_generator_set_training(l2i_generator_t* g, int enable);

/* Data-gradient backward of the forward above (the walk-training gradient path):
 * given grad_image [B,3,size,size] fp32, writes grad_latent [B, n_latent, D] fp32.
 * Must follow a forward made in training mode with the same batch on the same stream.  No weight
 * gradients are produced (the generator is frozen); gradients w.r.t. the noise inputs are not needed
 * on the walk-training path and are not produced either. */
int l2i_generator_backward(l2i_generator_t* g, float* grad_latent, const float* grad_image,
                           int batch, void* stream);

/* Per-segment device timing of the last forward (CUDA events recorded on the forward's stream
 * around each layer's kernels).  profile_entry() waits for the segment's end event - it is the one
 * call of this library that blocks the host.  kind: 0 conv, 1 blur_act, 2 skip/rgb/image, 3 styles.
 * flops / bytes are the ALGORITHMIC figures of the reference formulation. */
int l2i_generator_set_profiling(l2i_generator_t* g, int enable);
int l2i_generator_profile_count(l2i_generator_t* g);
int l2i_generator_profile_entry(l2i_generator_t* g, int i, char* name, int name_len, int* kind, float* ms,
                                double* flops, double* bytes);

/* Debug / test taps: copies an internal activation of the last forward as fp32 NCHW.
 * name: "conv1", "convs.<j>", "skip.<k>" (k = 0 is to_rgb1).  out must hold the full tensor. */
int l2i_generator_read_activation(l2i_generator_t* g, const char* name, float* out, int64_t numel,
                                  int batch, void* stream);

/* image fp32 NCHW [B,3,H,W] -> uint8 NHWC, truncating (transform_base.py:551-552). */
int l2i_image_to_uint8(uint8_t* out, const float* image, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* L2I_B200_H */
