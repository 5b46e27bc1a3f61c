// Halo-tile 3x3 convolution plans (conv_halo.cu), used by dslb_conv_plan_* for single-segment narrow 3x3 convs.
#pragma once
#include "common.h"

struct dslb_halo_plan;

namespace dslb {
bool halo_eligible(const dslb_conv_seg_t& s);
int halo_plan_create(const dslb_conv_seg_t& s, dslb_halo_plan** out);
int halo_plan_run(const dslb_halo_plan* h, void* stream);
void halo_plan_destroy(dslb_halo_plan* h);
}  // namespace dslb
