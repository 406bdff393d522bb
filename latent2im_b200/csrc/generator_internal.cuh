// Internal definitions shared by generator.cu (forward) and backward.cu (data-gradient backward).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "conv_common.cuh"

namespace l2i {

// kernels implemented in the other translation units
template <typename T> int launch_conv_simt(const void*, const float*, const ConvGeom&, const EpiParams&, cudaStream_t);
int launch_conv_tc(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st);
bool conv_tc_supported(const ConvGeom& g, const EpiParams& e);
int conv_tc_block_n(const ConvGeom& g);
bool conv_tc_halo_supported(const ConvGeom& g, const EpiParams& e);
int launch_conv_tc_halo(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st);
int launch_pack_pair_weight(__nv_bfloat16* dst, const float* src, int Cout, float scale, cudaStream_t st);
bool conv_tc_ares_supported(const ConvGeom& g, const EpiParams& e);
int launch_conv_tc_ares(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st);
bool conv_tc_hring_supported(const ConvGeom& g, const EpiParams& e);
int launch_conv_tc_hring(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st);
bool conv_tc_vpair_supported(const ConvGeom& g, const EpiParams& e);
int launch_conv_tc_vpair(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st);
int launch_pack_vpair_weight(__nv_bfloat16* dst, const float* src, float scale, cudaStream_t st);
bool conv_tc_quad_supported(const ConvGeom& g, const EpiParams& e);
int launch_conv_tc_quad(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st);
int launch_pack_quad_weight(__nv_bfloat16* dst, const float* src, float scale, cudaStream_t st);
bool conv_tc_uprow_supported(const ConvGeom& g, const EpiParams& e);
int launch_conv_tc_uprow(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st);
int64_t uprow_weight_elems(int Cin, int Cout);
int launch_pack_uprow_weight(__nv_bfloat16* dst, const float* src, int Cout, int Cin, float scale, const float* fir, cudaStream_t st);
int launch_pack_composite_weight(__nv_bfloat16* dst, const float* src, int Cout, int Cin, float scale, const float* fir, cudaStream_t st);
template <typename T, typename TIN>
int launch_blur_act(void*, void*, const void*, int, int, int, int, int, int, const float*, int64_t, const float*,
                    const float*, const float*, int64_t, const float*, int, cudaStream_t);
bool fir_tma_supported(int C);
int launch_blur_act_tma(void*, void*, const void*, int, int, int, int, int, int, const float*, int64_t, const float*, const float*,
                        const float*, int64_t, const float*, int, cudaStream_t);
int launch_blur_bwd_tma(void*, const void*, const void*, const float*, int64_t, float*, int64_t, int, int, int, int, int, int,
                        const float*, cudaStream_t);
int launch_skip_combine(float*, const float*, int, const float*, const float*, int, int, int, const float*, cudaStream_t);
template <typename T> int launch_const_input(void*, const float*, const float*, int64_t, int, int, int, cudaStream_t);
int launch_demod(float*, int64_t, const float*, int64_t, const float*, const int64_t*, const int*, const int*, int, int, cudaStream_t);
int launch_rgb_weight(float*, int64_t, const float*, const int*, const float*, int64_t, int, int, cudaStream_t);
int launch_gather_latent(float*, const float*, int64_t, int64_t, int, int, int, cudaStream_t);
int launch_pack_conv_weight(float*, __nv_bfloat16*, float*, const float*, int, int, int, float, int, cudaStream_t);
int launch_scale_copy(float*, const float*, int64_t, float, cudaStream_t);
template <typename T> int launch_nhwc_to_nchw(float*, const void*, int, int, int, int, const float*, int64_t, cudaStream_t);
int launch_mapping_fused(float*, const float*, const float* const*, const float* const*, int, int, int, float, float, cudaStream_t);
int launch_linear(float*, int64_t, const float*, int64_t, const int*, const float*, const float*, int, int, int,
                  float, float, int, float, float, cudaStream_t);

struct Param {
  float* ptr = nullptr;
  int64_t numel = 0;
  bool set = false;
};

struct StyledConvLayer {
  std::string name;  // "conv1" or "convs.j"
  int cin, cout, res_in, res_out;
  bool up;
  int latent_idx, noise_idx;
  int s_off;   // offset of this layer's styles inside a row of s_all
  int d_off;   // offset of this layer's demod coefficients inside a row of d_all
  float* w_f32 = nullptr;            // [9][Cin][Cout]
  __nv_bfloat16* w_bf16 = nullptr;   // [9][Cout][Cin], or [18][Cout][Cin] = bf16 hi halves then lo residuals (split)
  __nv_bfloat16* w_pair = nullptr;   // Cin == 32 plain layers: [12][Cout][64] pair-packed tiles for the halo kernel
  __nv_bfloat16* w_vpair = nullptr;  // 64 -> 64 plain layers: [kw][2-kh][64][64] tiles for the vertical-pair kernel (conv_tc_vpair.cu)
  __nv_bfloat16* w_quad = nullptr;   // 32 -> 32 plain layers: [8][128][64] 2x2-block weight matrix (conv_tc_quad.cu)
  __nv_bfloat16* w_comp = nullptr;   // composite up-conv (transposed conv + blur folded): [9][4*Cout][Cin], rows (phase, co)
  __nv_bfloat16* w_uprow = nullptr;  // row-marching fused up-conv (conv_tc_uprow.cu): [Cout/CO][9 = (kh, dx)][2*CO = (b, co)][Cin], horizontal blur folded
  bool composite = false;            // inference forward of this up layer runs the composite kernel (no t intermediate)
  bool split = false;                // tensor-core path uses hi + lo weights (K doubled) to remove the weight rounding error
  float* w_f32_t = nullptr;          // [9][Cout][Cin] fp32, data-gradient convs (training only)
  __nv_bfloat16* w_bf16_t = nullptr; // [9][Cin][Cout] bf16: data-gradient conv on the tcgen05 kernel (GEMM N = Cin, K = Cout)
  // training state: saved activation y, saved raw up-conv output t, noise used by the last forward
  void* y_save = nullptr;
  void* t_save = nullptr;
  const float* noise_ptr = nullptr;
  int64_t noise_bs = 0;
  int64_t wsq_off = 0;
};

struct RgbLayer {
  std::string name;  // "to_rgb1" or "to_rgbs.k"
  int cin, res, latent_idx;
  int s_off;    // inside s_all row
  int wr_off;   // inside wr_all row
};

}  // namespace l2i

using namespace l2i;

struct l2i_generator {
  int size, D, n_mlp, cm, dtype, max_batch, log_size, num_layers, n_latent;
  float lr_mlp;
  float fir[4];  // flipped separable taps * 2 (== taps / sum * 2)
  bool finalized = false;
  bool training = false;        // forward keeps what backward needs
  bool train_buffers = false;   // buffers below are allocated
  bool train_weights_packed = false;  // data-gradient weight copies match the current parameters (reset by finalize)
  int last_batch = 0;
  int last_train_batch = 0;
  void* gbuf = nullptr;         // gradient scratch (activation sized)
  float *R_s = nullptr, *R_d = nullptr, *R_rgb = nullptr, *gs_all = nullptr;  // reductions / style grads
  float* gskip[2] = {nullptr, nullptr};
  float* R_s0 = nullptr;        // style-gradient reduction of conv1 (through the constant input)
  float* fir2d_dev = nullptr;   // 4x4 FIR of the skip up-sampling, flipped, for the transposed op
  int* lat_seg = nullptr;       // [n_latent][1 + 2*3]: count, (row_start, row_count) x 3
  int conv_impl = 0;  // 0 auto, 1 simt, 2 tc
  // layers with res_out <= this use split-bf16 (hi + lo) weights on the tensor-core path: 16 keeps the tiny low-resolution
  // layers exact at no measurable cost; 64 buys ~2.5 dB more PSNR for ~0.9 ms per 32-image step (L2I_SPLIT_RES)
  int split_max_res = 16;
  int composite_min_res = 256;  // up layers with res_out >= this fold the blur into the conv weights (bf16 inference path)

  std::unordered_map<std::string, Param> params;
  std::vector<StyledConvLayer> convs;
  std::vector<RgbLayer> rgbs;

  // style tables
  int s_rows = 0, d_rows = 0, wr_elems = 0;
  float *mod_w_all = nullptr, *mod_b_all = nullptr;
  int* row_xoff = nullptr;
  float* wsq_all = nullptr;
  int64_t* row_wsq_off = nullptr;
  int *row_s_off = nullptr, *row_cin = nullptr;
  float* wrgb_all = nullptr;
  int* rgb_elem_s_off = nullptr;

  // workspace
  float *latent_buf = nullptr, *s_all = nullptr, *d_all = nullptr, *wr_all = nullptr;
  void* act[2] = {nullptr, nullptr};
  void* tbuf = nullptr;
  float* rgb_part = nullptr;
  float* skip[2] = {nullptr, nullptr};
  float* map_buf[2] = {nullptr, nullptr};
  const float** map_w_ptrs = nullptr;   // device tables of the mapping layers' weight / bias pointers (fused mapping kernel)
  const float** map_b_ptrs = nullptr;
  // where each layer's output landed in the last forward (debug taps)
  std::vector<const void*> conv_out;
  std::vector<const float*> skip_out;

  std::vector<void*> allocs;
  size_t elem_size() const { return dtype == L2I_F32 ? 4 : 2; }

  // optional per-segment timing (CUDA events on the caller's stream)
  struct Segment {
    std::string name;
    int kind;          // 0 conv (tensor / FFMA bound), 1 blur_act, 2 skip / rgb, 3 styles & misc
    double flops, bytes;
    cudaEvent_t ev0, ev1;
  };
  bool profiling = false;
  std::vector<Segment> segs;
  size_t seg_used = 0;
  Segment* seg_begin(const std::string& name, int kind, double flops, double bytes, cudaStream_t st) {
    if (!profiling) return nullptr;
    if (seg_used == segs.size()) {
      Segment sg;
      cudaEventCreate(&sg.ev0);
      cudaEventCreate(&sg.ev1);
      segs.push_back(sg);
    }
    Segment& sg = segs[seg_used++];
    sg.name = name; sg.kind = kind; sg.flops = flops; sg.bytes = bytes;
    cudaEventRecord(sg.ev0, st);
    return &sg;
  }
  void seg_end(Segment* sg, cudaStream_t st) {
    if (sg) cudaEventRecord(sg->ev1, st);
  }
};

namespace l2i {

template <typename T>
inline int train_alloc(l2i_generator* g, T** p, int64_t n) {
  *p = nullptr;
  if (n <= 0) return L2I_OK;
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, (size_t)n * sizeof(T));
  if (e != cudaSuccess) {
    set_error("generator: cudaMalloc(%lld bytes) failed: %s", (long long)(n * (int64_t)sizeof(T)), cudaGetErrorString(e));
    return L2I_ERR_CUDA;
  }
  g->allocs.push_back(q);
  *p = (T*)q;
  return L2I_OK;
}

inline float* P(l2i_generator* g, const std::string& key) { return g->params.at(key).ptr; }
inline float* P_noise_w(l2i_generator* g, const StyledConvLayer& L) { return P(g, L.name + ".noise.weight"); }
inline float* P_bias(l2i_generator* g, const StyledConvLayer& L) { return P(g, L.name + ".activate.bias"); }

}  // namespace l2i
