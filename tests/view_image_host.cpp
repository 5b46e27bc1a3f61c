// TEST INFRASTRUCTURE ONLY — compiles dsl_b200/csrc/view_image.cuh (the per-pixel arithmetic view_images_kernel runs on
// the GPU) with g++ and loops it over one image on the host, so tests/test_image_oracle.py can pin that arithmetic
// against oracle/image_oracle.py and the reference golden on the GPU-less build box. Not part of libdslb.so.
#include "../dsl_b200/csrc/view_image.cuh"

extern "C" void view_image_host(const uint8_t* src, const int32_t* view8, const float* mean, const float* stdv, int to_rgb,
                                float* out, int H, int W) {
  dslb::ImageViewDev v;
  v.src_h = view8[0]; v.src_w = view8[1]; v.img_h = view8[2]; v.img_w = view8[3];
  v.ps_mode = view8[4]; v.ps_crop = view8[5]; v.flip = view8[6]; v.reserved = 0;
  double inv[3];
  for (int c = 0; c < 3; ++c) inv[c] = 1.0 / (double)stdv[c];
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float o[3] = {0.f, 0.f, 0.f};
      if (y < v.img_h && x < v.img_w) dslb::vi_pixel(src, v, y, x, mean, inv, to_rgb, o);
      for (int c = 0; c < 3; ++c) out[((long long)c * H + y) * W + x] = o[c];
    }
}

// The whole launch as dslb_view_images issues it: grid (ceil(W/32), ceil(H/32), B), block (32, 8), every thread through
// the same vi_thread the __global__ kernel calls. `out` should be pre-filled with NaN by the caller to prove full coverage.
extern "C" void view_images_grid_host(const uint8_t* const* srcs, const int32_t* views8, int B, const float* mean,
                                      const float* stdv, int to_rgb, float* out, int H, int W) {
  dslb::ViewImageParams prm;
  for (int c = 0; c < 3; ++c) { prm.mean[c] = mean[c]; prm.inv_std[c] = 1.0 / (double)stdv[c]; }
  prm.to_rgb = to_rgb ? 1 : 0;
  const dslb::ImageViewDev* views = reinterpret_cast<const dslb::ImageViewDev*>(views8);
  const int gx = (W + dslb::VI_TX - 1) / dslb::VI_TX, gy = (H + dslb::VI_TY * dslb::VI_ROWS - 1) / (dslb::VI_TY * dslb::VI_ROWS);
  for (int bz = 0; bz < B; ++bz)
    for (int by = 0; by < gy; ++by)
      for (int bx = 0; bx < gx; ++bx)
        for (int ty = 0; ty < dslb::VI_TY; ++ty)
          for (int tx = 0; tx < dslb::VI_TX; ++tx) dslb::vi_thread(srcs, views, prm, out, H, W, bx, by, bz, tx, ty);
}
