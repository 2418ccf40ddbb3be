// Runs the DEVICE code of torch_scae_b200/csrc/conv_cols.cu on the CPU (tests/emu/simt.h).
//   conv_cols_emu <in.bin> <out.bin>     in: B C H W stride group | x[B,C,H,W] | dcols[B*L, C*9]    out: cols | dx
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "simt.h"
#include "scae_b200.h"

namespace scae {
#include "common_device.inc"
#include "conv_cols_device.inc"
}  // namespace scae

int main(int argc, char** argv) {
  using namespace scae;
  if (argc != 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int h[6];
  if (fread(h, 4, 6, f) != 6) return 2;
  const int B = h[0], C = h[1], H = h[2], W = h[3], stride = h[4];
  const int Ho = (H - 3) / stride + 1, Wo = (W - 3) / stride + 1, L = Ho * Wo;
  std::vector<float> x((size_t)B * C * H * W), dcols((size_t)B * L * C * 9);
  if (fread(x.data(), 4, x.size(), f) != x.size() || fread(dcols.data(), 4, dcols.size(), f) != dcols.size()) return 2;
  fclose(f);
  std::vector<float> cols(dcols.size(), -7.f), dx(x.size(), -7.f);
  const int groups = (C + kColsGroup - 1) / kColsGroup;
  const int group2 = h[5];               // channel group of the col2im instantiation under test: 16 or 32
  const int groups2 = (C + group2 - 1) / group2;
  emu_launch(dim3(groups, B), dim3(kColsThreads), [&] { im2col3x3_kernel(x.data(), cols.data(), C, H, W, Ho, Wo, stride); });
  emu_launch(dim3(groups2, B), dim3(kColsThreads), [&] {
    if (stride == 2 && group2 == 16) col2im3x3_kernel<2, 16>(dcols.data(), dx.data(), C, H, W, Ho, Wo);
    else if (stride == 2) col2im3x3_kernel<2, kColsGroup>(dcols.data(), dx.data(), C, H, W, Ho, Wo);
    else if (group2 == 16) col2im3x3_kernel<1, 16>(dcols.data(), dx.data(), C, H, W, Ho, Wo);
    else col2im3x3_kernel<1, kColsGroup>(dcols.data(), dx.data(), C, H, W, Ho, Wo);
  });
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 1;
  fwrite(cols.data(), 4, cols.size(), o);
  fwrite(dx.data(), 4, dx.size(), o);
  fclose(o);
  return 0;
}
