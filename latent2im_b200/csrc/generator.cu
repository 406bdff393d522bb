// Whole-network launcher for the StyleGAN2 synthesis forward (reference networks.py:360-514) over
// the fused kernels of this library.  Owns the packed weights, the style tables and the activation
// workspace; forward() only enqueues kernels on the caller's stream.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "generator_internal.cuh"

namespace {

int channels_at(int res, int cm) {
  switch (res) {
    case 4: case 8: case 16: case 32: return 512;
    case 64: return 256 * cm;
    case 128: return 128 * cm;
    case 256: return 64 * cm;
    case 512: return 32 * cm;
    case 1024: return 16 * cm;
    default: return -1;
  }
}

template <typename T>
int dev_alloc(l2i_generator* g, T** p, int64_t n) { return train_alloc(g, p, n); }

int add_param(l2i_generator* g, const std::string& key, int64_t numel) {
  Param p;
  p.numel = numel;
  L2I_TRY(dev_alloc(g, &p.ptr, numel));
  g->params[key] = p;
  return L2I_OK;
}

TapList plain_taps() {
  TapList t{};
  t.n = 9;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      int i = kh * 3 + kw;
      t.dy[i] = (int8_t)(kh - 1);
      t.dx[i] = (int8_t)(kw - 1);
      t.wtap[i] = (int8_t)i;
    }
  return t;
}

// stride-2 transposed 3x3 conv, output phase (py, px): t[2*oy+py, 2*ox+px] gathers the taps kh = py
// (mod 2) from input row oy - kh/2 (conv_transpose2d: u = 2*y + kh), same for columns.
TapList upconv_taps(int py, int px) {
  TapList t{};
  t.n = 0;
  for (int kh = py; kh < 3; kh += 2)
    for (int kw = px; kw < 3; kw += 2) {
      t.dy[t.n] = (int8_t)(-(kh / 2));
      t.dx[t.n] = (int8_t)(-(kw / 2));
      t.wtap[t.n] = (int8_t)(kh * 3 + kw);
      ++t.n;
    }
  return t;
}

// Plain 32 -> 32 layers on even images run on the 2x2-block kernel (plain NHWC input).
bool layer_uses_quad(l2i_generator* g, const StyledConvLayer& L, int B) {
  if (g->dtype != L2I_BF16 || g->conv_impl == 1 || L.up || L.split || L.w_quad == nullptr) return false;
  ConvGeom geom{};
  geom.B = B; geom.H = geom.W = L.res_in; geom.Cin = L.cin; geom.Cout = L.cout; geom.in_scale = 1; geom.weight_taps = 9;
  geom.OH = geom.OW = L.res_in; geom.nphase = 1; geom.out_scale = 1; geom.out_H = geom.out_W = L.res_out;
  EpiParams e{};
  e.wr = g->wr_all;
  e.fused_skip = 1;
  return conv_tc_quad_supported(geom, e);
}

// Otherwise plain Cin == 32 layers run on the halo kernel with vertically pair-packed input units; their producer
// (the blur of the preceding up-conv) then writes that layout directly.
bool layer_uses_pair_halo(l2i_generator* g, const StyledConvLayer& L, int B) {
  if (g->dtype != L2I_BF16 || g->conv_impl == 1 || L.up || L.split || L.cin != 32 || L.w_pair == nullptr) return false;
  if (layer_uses_quad(g, L, B)) return false;
  ConvGeom geom{};
  geom.B = B; geom.H = geom.W = L.res_in; geom.Cin = L.cin; geom.Cout = L.cout; geom.in_scale = 1; geom.weight_taps = 9;
  geom.OH = geom.OW = L.res_in; geom.nphase = 1; geom.out_scale = 1; geom.out_H = geom.out_W = L.res_out;
  EpiParams e{};
  e.wr = g->wr_all;  // plain layers always carry their ToRGB / rgb-partial epilogue
  return conv_tc_halo_supported(geom, e);
}

int run_conv(l2i_generator* g, const StyledConvLayer& L, const void* in, const ConvGeom& geom, const EpiParams& e,
             cudaStream_t st) {
  if (g->dtype == L2I_F32) return launch_conv_simt<float>(in, L.w_f32, geom, e, st);
  const bool want_tc = g->conv_impl != 1;
  if (want_tc && !L.split && L.w_quad != nullptr && !geom.in_pair_packed && conv_tc_quad_supported(geom, e))
    return launch_conv_tc_quad(in, L.w_quad, geom, e, st);
  if (want_tc && !L.split && conv_tc_ares_supported(geom, e)) return launch_conv_tc_ares(in, L.w_bf16, geom, e, st);
  if (want_tc && !L.split && L.w_vpair != nullptr && conv_tc_vpair_supported(geom, e)) return launch_conv_tc_vpair(in, L.w_vpair, geom, e, st);
  if (want_tc && !L.split && conv_tc_halo_supported(geom, e) && (geom.Cin != 32 || (L.w_pair != nullptr && geom.in_pair_packed)))
    return launch_conv_tc_halo(in, geom.Cin == 32 ? L.w_pair : L.w_bf16, geom, e, st);
  if (want_tc && !L.split && conv_tc_hring_supported(geom, e)) return launch_conv_tc_hring(in, L.w_bf16, geom, e, st);
  if (want_tc && conv_tc_supported(geom, e)) {
    if (!L.split) return launch_conv_tc(in, L.w_bf16, geom, e, st);
    // split-bf16 weights: every tap is issued twice, against the hi and the lo half of the weight tensor
    ConvGeom g2 = geom;
    g2.weight_taps = 18;
    for (int ph = 0; ph < geom.nphase; ++ph) {
      TapList& t = g2.taps[ph];
      const int n = t.n;
      for (int i = 0; i < n; ++i) { t.dy[n + i] = t.dy[i]; t.dx[n + i] = t.dx[i]; t.wtap[n + i] = (int8_t)(t.wtap[i] + 9); }
      t.n = 2 * n;
    }
    return launch_conv_tc(in, L.w_bf16, g2, e, st);
  }
  if (g->conv_impl == 2) {
    set_error("generator: L2I_CONV_IMPL=tc but layer %s is not supported by the tcgen05 kernel", L.name.c_str());
    return L2I_ERR_UNSUPPORTED;
  }
  // no tcgen05 variant takes this layer shape: say so (once per layer) instead of silently running ~20x slower on CUDA cores
  static std::unordered_map<std::string, bool> warned;
  if (!warned[L.name]) {
    warned[L.name] = true;
    std::fprintf(stderr, "l2i_b200: warning: bf16 layer %s (Cin %d, Cout %d, %dx%d -> %d) falls back to the CUDA-core conv kernel; "
                         "set L2I_CONV_IMPL=tc to make this an error\n", L.name.c_str(), L.cin, L.cout, L.res_in, L.res_in, L.res_out);
  }
  return launch_conv_simt<__nv_bfloat16>(in, L.w_f32, geom, e, st);
}

}  // namespace

extern "C" int l2i_generator_create(l2i_generator_t** out, int size, int style_dim, int n_mlp, int channel_multiplier,
                                    const float* blur_taps, int n_blur_taps, float lr_mlp, int dtype, int max_batch) {
  L2I_REQUIRE(out != nullptr, "generator_create: null out");
  *out = nullptr;
  L2I_REQUIRE(size >= 8 && size <= 1024 && (size & (size - 1)) == 0, "generator_create: size %d must be a power of two in 8..1024", size);
  L2I_REQUIRE(style_dim >= 1 && n_mlp >= 0 && channel_multiplier >= 1, "generator_create: bad style_dim/n_mlp/channel_multiplier");
  L2I_REQUIRE(dtype == L2I_F32 || dtype == L2I_BF16, "generator_create: dtype must be L2I_F32 or L2I_BF16");
  L2I_REQUIRE(max_batch >= 1, "generator_create: max_batch must be >= 1");
  L2I_REQUIRE(blur_taps != nullptr && n_blur_taps == 4,
              "generator_create: the fused up-conv/blur path needs a 4-tap separable blur kernel (got %d taps)", n_blur_taps);

  refresh_kernel_switches();
  l2i_generator* g = new l2i_generator();
  g->size = size; g->D = style_dim; g->n_mlp = n_mlp; g->cm = channel_multiplier; g->dtype = dtype;
  g->max_batch = max_batch; g->lr_mlp = lr_mlp;
  g->log_size = (int)std::lround(std::log2((double)size));
  g->num_layers = (g->log_size - 2) * 2 + 1;
  g->n_latent = g->log_size * 2 - 2;
  {
    float sum = 0.f;
    for (int i = 0; i < 4; ++i) sum += blur_taps[i];
    // make_kernel(k) * factor^2 == outer(f, f) with f = taps / sum * 2; upfirdn2d correlates with the flip
    for (int i = 0; i < 4; ++i) g->fir[i] = blur_taps[3 - i] / sum * 2.f;
  }
  if (const char* env = std::getenv("L2I_CONV_IMPL")) {
    if (!std::strcmp(env, "simt")) g->conv_impl = 1;
    else if (!std::strcmp(env, "tc")) g->conv_impl = 2;
  }

  if (const char* env = std::getenv("L2I_SPLIT_RES")) g->split_max_res = std::atoi(env);
  if (const char* env = std::getenv("L2I_COMPOSITE_RES")) g->composite_min_res = std::atoi(env);
  int rc = L2I_OK;
  auto fail = [&](int code) { l2i_generator_destroy(g); return code; };
  const int D = style_dim;

  // ---- layer table (networks.py:396-438) ----
  {
    StyledConvLayer c1;
    c1.name = "conv1"; c1.cin = c1.cout = channels_at(4, g->cm); c1.res_in = c1.res_out = 4; c1.up = false;
    c1.latent_idx = 0; c1.noise_idx = 0;
    g->convs.push_back(c1);
    RgbLayer r1;
    r1.name = "to_rgb1"; r1.cin = c1.cout; r1.res = 4; r1.latent_idx = 1;
    g->rgbs.push_back(r1);
    int cin = c1.cout;
    for (int k = 0; k < g->log_size - 2; ++k) {
      const int res_in = 4 << k, res_out = res_in * 2;
      const int cout = channels_at(res_out, g->cm);
      StyledConvLayer up;
      up.name = "convs." + std::to_string(2 * k); up.cin = cin; up.cout = cout; up.res_in = res_in; up.res_out = res_out;
      up.up = true; up.latent_idx = 2 * k + 1; up.noise_idx = 2 * k + 1;
      g->convs.push_back(up);
      StyledConvLayer cv;
      cv.name = "convs." + std::to_string(2 * k + 1); cv.cin = cout; cv.cout = cout; cv.res_in = cv.res_out = res_out;
      cv.up = false; cv.latent_idx = 2 * k + 2; cv.noise_idx = 2 * k + 2;
      g->convs.push_back(cv);
      RgbLayer r;
      r.name = "to_rgbs." + std::to_string(k); r.cin = cout; r.res = res_out; r.latent_idx = 2 * k + 3;
      g->rgbs.push_back(r);
      cin = cout;
    }
  }

  // ---- parameter slots (rosinality state_dict keys) ----
  for (int i = 1; i <= n_mlp && rc == L2I_OK; ++i) {
    rc = add_param(g, "style." + std::to_string(i) + ".weight", (int64_t)D * D);
    if (rc == L2I_OK) rc = add_param(g, "style." + std::to_string(i) + ".bias", D);
  }
  if (rc == L2I_OK) rc = add_param(g, "input.input", (int64_t)g->convs[0].cin * 16);
  for (auto& L : g->convs) {
    if (rc != L2I_OK) break;
    rc = add_param(g, L.name + ".conv.weight", (int64_t)L.cout * L.cin * 9);
    if (rc == L2I_OK) rc = add_param(g, L.name + ".conv.modulation.weight", (int64_t)L.cin * D);
    if (rc == L2I_OK) rc = add_param(g, L.name + ".conv.modulation.bias", L.cin);
    if (rc == L2I_OK) rc = add_param(g, L.name + ".noise.weight", 1);
    if (rc == L2I_OK) rc = add_param(g, L.name + ".activate.bias", L.cout);
  }
  for (auto& R : g->rgbs) {
    if (rc != L2I_OK) break;
    rc = add_param(g, R.name + ".bias", 3);
    if (rc == L2I_OK) rc = add_param(g, R.name + ".conv.weight", (int64_t)3 * R.cin);
    if (rc == L2I_OK) rc = add_param(g, R.name + ".conv.modulation.weight", (int64_t)R.cin * D);
    if (rc == L2I_OK) rc = add_param(g, R.name + ".conv.modulation.bias", R.cin);
  }
  if (rc != L2I_OK) return fail(rc);

  // ---- style / demod / rgb tables ----
  int s_rows = 0, d_rows = 0, wr_elems = 0;
  int64_t wsq_total = 0;
  for (auto& L : g->convs) { L.s_off = s_rows; s_rows += L.cin; L.d_off = d_rows; d_rows += L.cout; L.wsq_off = wsq_total; wsq_total += (int64_t)L.cout * L.cin; }
  for (auto& R : g->rgbs) { R.s_off = s_rows; s_rows += R.cin; R.wr_off = wr_elems; wr_elems += 3 * R.cin; }
  g->s_rows = s_rows; g->d_rows = d_rows; g->wr_elems = wr_elems;
  rc = dev_alloc(g, &g->mod_w_all, (int64_t)s_rows * D);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->mod_b_all, s_rows);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->row_xoff, s_rows);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->wsq_all, wsq_total);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->row_wsq_off, d_rows);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->row_s_off, d_rows);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->row_cin, d_rows);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->wrgb_all, wr_elems);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->rgb_elem_s_off, wr_elems);
  if (rc != L2I_OK) return fail(rc);
  {
    std::vector<int> xoff(s_rows), rs(d_rows), rc_(d_rows), eoff(wr_elems);
    std::vector<int64_t> rw(d_rows);
    for (auto& L : g->convs) {
      for (int i = 0; i < L.cin; ++i) xoff[L.s_off + i] = L.latent_idx * D;
      for (int co = 0; co < L.cout; ++co) {
        rw[L.d_off + co] = L.wsq_off + (int64_t)co * L.cin;
        rs[L.d_off + co] = L.s_off;
        rc_[L.d_off + co] = L.cin;
      }
    }
    for (auto& R : g->rgbs) {
      for (int i = 0; i < R.cin; ++i) xoff[R.s_off + i] = R.latent_idx * D;
      for (int c = 0; c < 3; ++c)
        for (int i = 0; i < R.cin; ++i) eoff[R.wr_off + c * R.cin + i] = R.s_off + i;
    }
    cudaError_t e = cudaMemcpy(g->row_xoff, xoff.data(), sizeof(int) * s_rows, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->row_wsq_off, rw.data(), sizeof(int64_t) * d_rows, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->row_s_off, rs.data(), sizeof(int) * d_rows, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->row_cin, rc_.data(), sizeof(int) * d_rows, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->rgb_elem_s_off, eoff.data(), sizeof(int) * wr_elems, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("generator_create: table upload failed: %s", cudaGetErrorString(e)); return fail(L2I_ERR_CUDA); }
  }

  // ---- packed weights ----
  for (auto& L : g->convs) {
    rc = dev_alloc(g, &L.w_f32, (int64_t)9 * L.cin * L.cout);
    L.split = dtype == L2I_BF16 && L.res_out <= g->split_max_res;
    if (rc == L2I_OK && dtype == L2I_BF16) rc = dev_alloc(g, &L.w_bf16, (int64_t)(L.split ? 18 : 9) * L.cin * L.cout);
    if (rc == L2I_OK && dtype == L2I_BF16 && L.cin == 32 && !L.up) rc = dev_alloc(g, &L.w_pair, (int64_t)12 * L.cout * 64);
    if (rc == L2I_OK && dtype == L2I_BF16 && L.cin == 32 && L.cout == 32 && !L.up) rc = dev_alloc(g, &L.w_quad, (int64_t)128 * 512);
    if (rc == L2I_OK && dtype == L2I_BF16 && L.cin == 64 && L.cout == 64 && !L.up) rc = dev_alloc(g, &L.w_vpair, (int64_t)9 * 64 * 64);
    L.composite = dtype == L2I_BF16 && L.up && g->conv_impl != 1 && L.res_out >= g->composite_min_res && L.cout % 32 == 0 &&
                  L.cin % 64 == 0;
    if (rc == L2I_OK && L.composite) rc = dev_alloc(g, &L.w_comp, (int64_t)36 * L.cin * L.cout);
    if (rc == L2I_OK && L.composite && uprow_weight_elems(L.cin, L.cout) > 0) rc = dev_alloc(g, &L.w_uprow, uprow_weight_elems(L.cin, L.cout));
    if (rc != L2I_OK) return fail(rc);
  }

  // ---- workspace ----
  const int64_t B = max_batch;
  int64_t act_elems = 0, t_elems = 0, part_elems = 0;
  for (auto& L : g->convs) {
    act_elems = std::max(act_elems, B * L.res_out * L.res_out * (int64_t)L.cout);
    act_elems = std::max(act_elems, B * L.res_in * L.res_in * (int64_t)L.cin);
    if (L.up) t_elems = std::max(t_elems, B * (int64_t)(2 * L.res_in + 2) * (2 * L.res_in + 2) * L.cout);
  }
  for (auto& R : g->rgbs) part_elems = std::max(part_elems, B * 3 * (int64_t)R.res * R.res * ceil_div(R.cin, 64));
  const size_t es = g->elem_size();
  {
    char *a0 = nullptr, *a1 = nullptr, *tb = nullptr;
    rc = dev_alloc(g, &a0, act_elems * (int64_t)es);
    if (rc == L2I_OK) rc = dev_alloc(g, &a1, act_elems * (int64_t)es);
    if (rc == L2I_OK) rc = dev_alloc(g, &tb, t_elems * (int64_t)es);
    g->act[0] = a0; g->act[1] = a1; g->tbuf = tb;
  }
  if (rc == L2I_OK) rc = dev_alloc(g, &g->rgb_part, part_elems);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->skip[0], B * 3 * (int64_t)size * size);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->skip[1], B * 3 * (int64_t)(size / 2) * (size / 2));
  if (rc == L2I_OK) rc = dev_alloc(g, &g->latent_buf, B * g->n_latent * (int64_t)D);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->s_all, B * (int64_t)s_rows);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->d_all, B * (int64_t)d_rows);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->wr_all, B * (int64_t)wr_elems);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->map_buf[0], B * (int64_t)D);
  if (rc == L2I_OK) rc = dev_alloc(g, &g->map_buf[1], B * (int64_t)D);
  if (rc != L2I_OK) return fail(rc);
  if (n_mlp > 0) {
    rc = dev_alloc(g, &g->map_w_ptrs, n_mlp);
    if (rc == L2I_OK) rc = dev_alloc(g, &g->map_b_ptrs, n_mlp);
    if (rc != L2I_OK) return fail(rc);
    std::vector<const float*> wp(n_mlp), bp(n_mlp);
    for (int i = 1; i <= n_mlp; ++i) { wp[i - 1] = P(g, "style." + std::to_string(i) + ".weight"); bp[i - 1] = P(g, "style." + std::to_string(i) + ".bias"); }
    cudaError_t e = cudaMemcpy(g->map_w_ptrs, wp.data(), sizeof(float*) * n_mlp, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g->map_b_ptrs, bp.data(), sizeof(float*) * n_mlp, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("generator_create: mapping table upload failed: %s", cudaGetErrorString(e)); return fail(L2I_ERR_CUDA); }
  }
  g->conv_out.assign(g->convs.size(), nullptr);
  g->skip_out.assign(g->rgbs.size(), nullptr);
  *out = g;
  return L2I_OK;
}

extern "C" void l2i_generator_destroy(l2i_generator_t* g) {
  if (!g) return;
  for (void* p : g->allocs) cudaFree(p);
  for (auto& sg : g->segs) { cudaEventDestroy(sg.ev0); cudaEventDestroy(sg.ev1); }
  delete g;
}

extern "C" int l2i_generator_num_layers(const l2i_generator_t* g) { return g ? g->num_layers : -1; }
extern "C" int l2i_generator_n_latent(const l2i_generator_t* g) { return g ? g->n_latent : -1; }

extern "C" int l2i_generator_set_param(l2i_generator_t* g, const char* key, const float* data, int64_t numel, void* stream) {
  L2I_REQUIRE(g && key, "generator_set_param: null argument");
  const std::string k(key);
  // buffers the native side does not consume: FIR kernels are fixed at create time, the registered
  // noise buffers are passed explicitly to forward()
  if (k.find(".blur.kernel") != std::string::npos || k.find(".upsample.kernel") != std::string::npos ||
      k.rfind("noises.", 0) == 0)
    return L2I_OK;
  auto it = g->params.find(k);
  L2I_REQUIRE(it != g->params.end(), "generator_set_param: unknown key '%s'", key);
  L2I_REQUIRE(it->second.numel == numel, "generator_set_param: '%s' expects %lld elements, got %lld", key,
              (long long)it->second.numel, (long long)numel);
  L2I_REQUIRE(data != nullptr, "generator_set_param: null data for '%s'", key);
  L2I_CUDA_TRY(cudaMemcpyAsync(it->second.ptr, data, sizeof(float) * numel, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  it->second.set = true;
  g->finalized = false;
  return L2I_OK;
}

extern "C" int l2i_generator_finalize(l2i_generator_t* g, void* stream) {
  L2I_REQUIRE(g, "generator_finalize: null generator");
  for (auto& kv : g->params)
    if (!kv.second.set) {
      set_error("generator_finalize: parameter '%s' was never set", kv.first.c_str());
      return L2I_ERR_STATE;
    }
  cudaStream_t st = (cudaStream_t)stream;
  const int D = g->D;
  const float mod_scale = 1.0f / std::sqrt((float)D);  // EqualLinear(style_dim, Cin, bias_init=1): lr_mul = 1
  for (auto& L : g->convs) {
    const float scale = 1.0f / std::sqrt((float)(L.cin * 9));
    L2I_TRY(launch_pack_conv_weight(L.w_f32, L.w_bf16, g->wsq_all + L.wsq_off, P(g, L.name + ".conv.weight"), L.cout,
                                    L.cin, 9, scale, L.split ? 1 : 0, st));
    if (L.w_pair) L2I_TRY(launch_pack_pair_weight(L.w_pair, P(g, L.name + ".conv.weight"), L.cout, scale, st));
    if (L.w_vpair) L2I_TRY(launch_pack_vpair_weight(L.w_vpair, P(g, L.name + ".conv.weight"), scale, st));
    if (L.w_quad) L2I_TRY(launch_pack_quad_weight(L.w_quad, P(g, L.name + ".conv.weight"), scale, st));
    if (L.w_comp) L2I_TRY(launch_pack_composite_weight(L.w_comp, P(g, L.name + ".conv.weight"), L.cout, L.cin, scale, g->fir, st));
    if (L.w_uprow) L2I_TRY(launch_pack_uprow_weight(L.w_uprow, P(g, L.name + ".conv.weight"), L.cout, L.cin, scale, g->fir, st));
    L2I_TRY(launch_scale_copy(g->mod_w_all + (int64_t)L.s_off * D, P(g, L.name + ".conv.modulation.weight"),
                              (int64_t)L.cin * D, mod_scale, st));
    L2I_TRY(launch_scale_copy(g->mod_b_all + L.s_off, P(g, L.name + ".conv.modulation.bias"), L.cin, 1.f, st));
  }
  for (auto& R : g->rgbs) {
    const float scale = 1.0f / std::sqrt((float)R.cin);
    L2I_TRY(launch_scale_copy(g->wrgb_all + R.wr_off, P(g, R.name + ".conv.weight"), (int64_t)3 * R.cin, scale, st));
    L2I_TRY(launch_scale_copy(g->mod_w_all + (int64_t)R.s_off * D, P(g, R.name + ".conv.modulation.weight"),
                              (int64_t)R.cin * D, mod_scale, st));
    L2I_TRY(launch_scale_copy(g->mod_b_all + R.s_off, P(g, R.name + ".conv.modulation.bias"), R.cin, 1.f, st));
  }
  g->finalized = true;
  g->train_weights_packed = false;
  return L2I_OK;
}

extern "C" int l2i_generator_mapping(l2i_generator_t* g, float* w, const float* z, int batch, void* stream) {
  L2I_REQUIRE(g && (batch == 0 || (w && z)), "generator_mapping: null argument");
  if (!g->finalized) { set_error("generator_mapping: call l2i_generator_finalize first"); return L2I_ERR_STATE; }
  if (batch > g->max_batch) { set_error("generator_mapping: batch %d > max_batch %d", batch, g->max_batch); return L2I_ERR_STATE; }
  if (batch == 0) return L2I_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = g->D;
  // PixelNorm + one batch-in-lanes linear launch per layer (weights read once per 32 latent rows; the single-launch fused kernel
  // streams the 8 MB of weights once per ROW and took 0.23 ms at batch 32).  L2I_MAPPING_FUSED=1 selects the fused kernel.
  static const bool fused_mapping = std::getenv("L2I_MAPPING_FUSED") != nullptr && std::atoi(std::getenv("L2I_MAPPING_FUSED")) != 0;
  if (fused_mapping && g->n_mlp > 0 && D % 4 == 0 && D <= 4096) {
    const float ws = (1.0f / std::sqrt((float)D)) * g->lr_mlp;
    return launch_mapping_fused(w, z, g->map_w_ptrs, g->map_b_ptrs, batch, g->n_mlp, D, ws, g->lr_mlp, st);
  }
  float* cur = g->n_mlp == 0 ? w : g->map_buf[0];
  L2I_TRY(l2i_pixel_norm(cur, z, batch, D, stream));
  const float wscale = (1.0f / std::sqrt((float)D)) * g->lr_mlp;
  for (int i = 1; i <= g->n_mlp; ++i) {
    float* nxt = (i == g->n_mlp) ? w : g->map_buf[i & 1];
    L2I_TRY(launch_linear(nxt, D, cur, D, nullptr, P(g, "style." + std::to_string(i) + ".weight"),
                          P(g, "style." + std::to_string(i) + ".bias"), batch, D, D, wscale, g->lr_mlp, 1, 0.2f,
                          1.4142135623730951f, st));
    cur = nxt;
  }
  return L2I_OK;
}

extern "C" int l2i_generator_forward(l2i_generator_t* g, const float* latent, int64_t latent_batch_stride,
                                     int64_t latent_layer_stride, const float* const* noise, const int* noise_batch,
                                     float* image, uint8_t* image_u8, int batch, void* stream) {
  L2I_REQUIRE(g, "generator_forward: null generator");
  if (!g->finalized) { set_error("generator_forward: call l2i_generator_finalize first"); return L2I_ERR_STATE; }
  if (batch > g->max_batch) { set_error("generator_forward: batch %d > max_batch %d", batch, g->max_batch); return L2I_ERR_STATE; }
  L2I_REQUIRE(batch >= 0, "generator_forward: negative batch");
  if (batch == 0) return L2I_OK;
  L2I_REQUIRE(latent != nullptr, "generator_forward: null latent");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = batch, D = g->D;
  const bool f32 = g->dtype == L2I_F32;
  for (int i = 0; i < g->num_layers; ++i)
    if (noise != nullptr && noise[i] != nullptr)
      L2I_REQUIRE(noise_batch != nullptr && (noise_batch[i] == 1 || noise_batch[i] == B),
                  "generator_forward: noise[%d] batch must be 1 or %d", i, B);

  g->seg_used = 0;
  const double es = (double)g->elem_size();
  // 1. styles for every modulated conv, demodulation coefficients, ToRGB effective weights
  auto* sg_styles = g->seg_begin("styles", 3, 2.0 * B * g->s_rows * D, 4.0 * g->s_rows * D, st);
  L2I_TRY(launch_gather_latent(g->latent_buf, latent, latent_batch_stride, latent_layer_stride, B, g->n_latent, D, st));
  L2I_TRY(launch_linear(g->s_all, g->s_rows, g->latent_buf, (int64_t)g->n_latent * D, g->row_xoff, g->mod_w_all,
                        g->mod_b_all, B, g->s_rows, D, 1.f, 1.f, 0, 0.f, 1.f, st));
  L2I_TRY(launch_demod(g->d_all, g->d_rows, g->s_all, g->s_rows, g->wsq_all, g->row_wsq_off, g->row_s_off, g->row_cin,
                       g->d_rows, B, st));
  L2I_TRY(launch_rgb_weight(g->wr_all, g->wr_elems, g->wrgb_all, g->rgb_elem_s_off, g->s_all, g->s_rows, g->wr_elems, B, st));

  // 2. constant input scaled by conv1's style
  int cur = 0;
  {
    const auto& L = g->convs[0];
    if (f32) L2I_TRY(launch_const_input<float>(g->act[cur], P(g, "input.input"), g->s_all + L.s_off, g->s_rows, B, L.cin, 16, st));
    else L2I_TRY(launch_const_input<__nv_bfloat16>(g->act[cur], P(g, "input.input"), g->s_all + L.s_off, g->s_rows, B, L.cin, 16, st));
  }
  g->seg_end(sg_styles, st);

  // 3. layers
  const float* skip_prev = nullptr;
  int skip_sel = (int)(g->rgbs.size() - 1) & 1;  // arrange that the final skip lands in skip[0] (full-size buffer)
  size_t rgb_i = 0;
  for (size_t li = 0; li < g->convs.size(); ++li) {
    const auto& L = g->convs[li];
    const bool last_conv = li + 1 == g->convs.size();
    const StyledConvLayer* next = last_conv ? nullptr : &g->convs[li + 1];
    const float* nz = noise ? noise[L.noise_idx] : nullptr;
    const int64_t nz_bs = (nz && noise_batch[L.noise_idx] == B && B > 1) ? (int64_t)L.res_out * L.res_out : 0;
    const float* nz_w = P(g, L.name + ".noise.weight");
    const float* s_next = next ? g->s_all + next->s_off : nullptr;

    ConvGeom geom{};
    geom.B = B; geom.H = geom.W = L.res_in; geom.Cin = L.cin; geom.Cout = L.cout; geom.in_scale = 1; geom.weight_taps = 9;
    const bool keep = g->training;
    if (keep) { g->convs[li].noise_ptr = nz; g->convs[li].noise_bs = nz_bs; }
    EpiParams e{};
    e.demod = g->d_all + L.d_off; e.demod_bs = g->d_rows;
    for (int i = 0; i < 4; ++i) e.fir[i] = g->fir[i];

    if (L.up && L.composite && !keep) {
      // transposed conv + blur + noise + bias + lrelu + next-style scale in ONE kernel: a 3x3 conv at input
      // resolution with N = 4 phases x Cout (composite 6x6 stride-2 kernel), no (2H+1)^2 intermediate
      geom.OH = geom.OW = L.res_in; geom.nphase = 1; geom.out_scale = 1; geom.out_H = geom.out_W = L.res_out;
      geom.taps[0] = plain_taps();
      geom.Cout = 4 * L.cout; geom.up_cout = L.cout;
      geom.out_pair_packed = (next != nullptr && layer_uses_pair_halo(g, *next, B)) ? 1 : 0;
      e.mode = 0;
      e.bias = P(g, L.name + ".activate.bias");
      e.noise = nz; e.noise_bs = nz_bs; e.noise_w = nz_w;
      e.s_next = s_next; e.s_next_bs = g->s_rows;
      e.out = g->act[cur ^ 1];
      e.rgb_part = g->rgb_part;   // scratch (only the cycle-accounting build of conv_tc_uprow.cu writes to it)
      const double px_in = (double)B * L.res_in * L.res_in, px_out = (double)B * L.res_out * L.res_out;
      auto* sg_c = g->seg_begin(L.name + "/upconv+blur_act", 0, 2.0 * 9 * L.cin * L.cout * px_in,
                                (px_in * L.cin + px_out * L.cout) * es + px_out * 4.0, st);
      if (L.w_uprow != nullptr && conv_tc_uprow_supported(geom, e)) L2I_TRY(launch_conv_tc_uprow(g->act[cur], L.w_uprow, geom, e, st));
      else if (conv_tc_halo_supported(geom, e)) L2I_TRY(launch_conv_tc_halo(g->act[cur], L.w_comp, geom, e, st));
      else if (conv_tc_ares_supported(geom, e)) L2I_TRY(launch_conv_tc_ares(g->act[cur], L.w_comp, geom, e, st));
      else if (conv_tc_supported(geom, e)) L2I_TRY(launch_conv_tc(g->act[cur], L.w_comp, geom, e, st));
      else { set_error("generator: composite up-conv of %s is not supported by the tcgen05 kernels", L.name.c_str()); return L2I_ERR_UNSUPPORTED; }
      g->seg_end(sg_c, st);
      cur ^= 1;
      g->conv_out[li] = g->act[cur];
      continue;
    }
    if (L.up) {
      geom.OH = geom.OW = L.res_in + 1; geom.nphase = 4; geom.out_scale = 2;
      geom.out_H = geom.out_W = 2 * L.res_in + 2;
      for (int ph = 0; ph < 4; ++ph) geom.taps[ph] = upconv_taps(ph >> 1, ph & 1);
      void* tdst = keep ? L.t_save : g->tbuf;
      e.mode = 1; e.out = tdst; e.raw_fp16 = f32 ? 0 : 1;
      const double px_in = (double)B * L.res_in * L.res_in, px_out = (double)B * L.res_out * L.res_out;
      const double px_t = (double)B * (2.0 * L.res_in + 1) * (2.0 * L.res_in + 1);
      auto* sg_c = g->seg_begin(L.name + "/upconv", 0, 2.0 * 9 * L.cin * L.cout * px_in,
                                (px_in * L.cin + px_t * L.cout) * es, st);
      L2I_TRY(run_conv(g, L, g->act[cur], geom, e, st));
      g->seg_end(sg_c, st);
      auto* sg_b = g->seg_begin(L.name + "/blur_act", 1, 0.0, (px_t + px_out) * L.cout * es + px_out * 4.0, st);
      void* dst = g->act[cur ^ 1];
      const bool pair_next = !f32 && next != nullptr && layer_uses_pair_halo(g, *next, B);
      if (f32) L2I_TRY(launch_blur_act<float, float>(dst, keep ? L.y_save : nullptr, tdst, B, L.res_out, L.res_out, L.cout, geom.out_H, geom.out_W, nz, nz_bs, nz_w,
                                              P(g, L.name + ".activate.bias"), s_next, g->s_rows, g->fir, 0, st));
      else if (fir_tma_supported(L.cout))
        L2I_TRY(launch_blur_act_tma(dst, keep ? L.y_save : nullptr, tdst, B, L.res_out, L.res_out, L.cout, geom.out_H, geom.out_W, nz, nz_bs,
                                    nz_w, P(g, L.name + ".activate.bias"), s_next, g->s_rows, g->fir, pair_next ? 1 : 0, st));
      else L2I_TRY(launch_blur_act<__nv_bfloat16, __half>(dst, keep ? L.y_save : nullptr, tdst, B, L.res_out, L.res_out, L.cout, geom.out_H, geom.out_W, nz, nz_bs,
                                                  nz_w, P(g, L.name + ".activate.bias"), s_next, g->s_rows, g->fir, pair_next ? 1 : 0, st));
      g->seg_end(sg_b, st);
      cur ^= 1;
      g->conv_out[li] = g->act[cur];
      continue;
    }

    // plain styled conv, followed by a ToRGB (conv1 -> to_rgb1, convs[odd] -> to_rgbs[k])
    const auto& R = g->rgbs[rgb_i];
    geom.OH = geom.OW = L.res_in; geom.nphase = 1; geom.out_scale = 1; geom.out_H = geom.out_W = L.res_out;
    geom.taps[0] = plain_taps();
    geom.in_pair_packed = (!f32 && layer_uses_pair_halo(g, L, B)) ? 1 : 0;
    e.mode = 0;
    e.bias = P(g, L.name + ".activate.bias");
    e.noise = nz; e.noise_bs = nz_bs; e.noise_w = nz_w;
    e.s_next = s_next; e.s_next_bs = g->s_rows;
    e.out = s_next ? g->act[cur ^ 1] : nullptr;
    e.y_out = keep ? L.y_save : nullptr;
    e.wr = g->wr_all + R.wr_off; e.wr_bs = g->wr_elems;
    e.rgb_bias = P(g, R.name + ".bias");
    e.skip_in = skip_prev;
    const bool final_rgb = rgb_i + 1 == g->rgbs.size();
    float* skip_dst = (final_rgb && image != nullptr) ? image : g->skip[skip_sel];
    e.skip_out = skip_dst;
    // last layer on the 2x2-block kernel with only the uint8 image wanted: the cast is fused into its epilogue
    bool fused_u8 = false;
    if (final_rgb && image == nullptr && image_u8 != nullptr && !keep && !f32 && layer_uses_quad(g, L, B)) {
      e.image_u8 = image_u8;
      e.skip_out = nullptr;
      fused_u8 = true;
    }
    e.rgb_part = g->rgb_part;
    const int n_tile = (f32 || g->conv_impl == 1 || !conv_tc_supported(geom, e)) ? 64 : conv_tc_block_n(geom);
    const int nparts = ceil_div(L.cout, n_tile);
    e.fused_skip = nparts == 1 ? 1 : 0;
    const double px = (double)B * L.res_out * L.res_out;
    auto* sg_c = g->seg_begin(L.name + "/conv+torgb", 0, 2.0 * (9.0 * L.cin + 3.0) * L.cout * px,
                              px * L.cin * es + (s_next ? px * L.cout * es : 0.0) + px * 4.0 + px * 12.0 * (e.fused_skip ? 1.25 : 1.0), st);
    L2I_TRY(run_conv(g, L, g->act[cur], geom, e, st));
    g->seg_end(sg_c, st);
    if (!e.fused_skip) {
      auto* sg_s = g->seg_begin(R.name + "/skip", 2, 0.0, px * 12.0 * (nparts + 1.25), st);
      L2I_TRY(launch_skip_combine(skip_dst, g->rgb_part, nparts, e.rgb_bias, skip_prev, B, L.res_out, L.res_out, g->fir, st));
      g->seg_end(sg_s, st);
    }
    if (s_next) cur ^= 1;
    g->conv_out[li] = s_next ? g->act[cur] : nullptr;
    g->skip_out[rgb_i] = fused_u8 ? nullptr : skip_dst;
    skip_prev = fused_u8 ? nullptr : skip_dst;
    skip_sel ^= 1;
    ++rgb_i;
  }
  if (image_u8 != nullptr && skip_prev != nullptr) {
    auto* sg_u = g->seg_begin("image_to_uint8", 2, 0.0, (double)B * g->size * g->size * 15.0, st);
    L2I_TRY(l2i_image_to_uint8(image_u8, skip_prev, B, g->size, g->size, stream));
    g->seg_end(sg_u, st);
  }
  g->last_batch = B;
  g->last_train_batch = g->training ? B : 0;   // any inference forward invalidates a pending backward (style tables / buffers reused)
  return L2I_OK;
}

extern "C" int l2i_generator_set_profiling(l2i_generator_t* g, int enable) {
  L2I_REQUIRE(g, "generator_set_profiling: null generator");
  g->profiling = enable != 0;
  g->seg_used = 0;
  return L2I_OK;
}

extern "C" int l2i_generator_profile_count(l2i_generator_t* g) { return g ? (int)g->seg_used : -1; }

extern "C" int l2i_generator_profile_entry(l2i_generator_t* g, int i, char* name, int name_len, int* kind, float* ms,
                                           double* flops, double* bytes) {
  L2I_REQUIRE(g && i >= 0 && (size_t)i < g->seg_used, "generator_profile_entry: index out of range");
  auto& sg = g->segs[i];
  L2I_CUDA_TRY(cudaEventSynchronize(sg.ev1));
  float t = 0.f;
  L2I_CUDA_TRY(cudaEventElapsedTime(&t, sg.ev0, sg.ev1));
  if (name && name_len > 0) { std::strncpy(name, sg.name.c_str(), name_len - 1); name[name_len - 1] = 0; }
  if (kind) *kind = sg.kind;
  if (ms) *ms = t;
  if (flops) *flops = sg.flops;
  if (bytes) *bytes = sg.bytes;
  return L2I_OK;
}

extern "C" int l2i_generator_read_activation(l2i_generator_t* g, const char* name, float* out, int64_t numel, int batch,
                                             void* stream) {
  L2I_REQUIRE(g && name && out, "generator_read_activation: null argument");
  if (batch != g->last_batch) { set_error("generator_read_activation: batch %d does not match the last forward (%d)", batch, g->last_batch); return L2I_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  const std::string n(name);
  if (n.rfind("skip.", 0) == 0) {
    const int k = std::atoi(n.c_str() + 5);
    L2I_REQUIRE(k >= 0 && k < (int)g->rgbs.size() && g->skip_out[k] != nullptr, "generator_read_activation: no such skip '%s'", name);
    const int64_t need = (int64_t)batch * 3 * g->rgbs[k].res * g->rgbs[k].res;
    L2I_REQUIRE(numel == need, "generator_read_activation: '%s' has %lld elements", name, (long long)need);
    L2I_CUDA_TRY(cudaMemcpyAsync(out, g->skip_out[k], sizeof(float) * need, cudaMemcpyDeviceToDevice, st));
    return L2I_OK;
  }
  for (size_t li = 0; li < g->convs.size(); ++li) {
    const auto& L = g->convs[li];
    if (L.name != n) continue;
    L2I_REQUIRE(g->conv_out[li] != nullptr && li + 1 < g->convs.size(), "generator_read_activation: '%s' is not materialised", name);
    const int64_t need = (int64_t)batch * L.cout * L.res_out * L.res_out;
    L2I_REQUIRE(numel == need, "generator_read_activation: '%s' has %lld elements", name, (long long)need);
    const float* inv = g->s_all + g->convs[li + 1].s_off;
    if (g->dtype == L2I_F32) return launch_nhwc_to_nchw<float>(out, g->conv_out[li], batch, L.res_out, L.res_out, L.cout, inv, g->s_rows, st);
    return launch_nhwc_to_nchw<__nv_bfloat16>(out, g->conv_out[li], batch, L.res_out, L.res_out, L.cout, inv, g->s_rows, st);
  }
  set_error("generator_read_activation: unknown activation '%s'", name);
  return L2I_ERR_INVALID_ARG;
}
