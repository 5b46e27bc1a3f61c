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
