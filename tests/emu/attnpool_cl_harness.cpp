// Runs the DEVICE code of torch_scae_b200/csrc/attnpool_cl.cu on the CPU (tests/emu/simt.h); the build script pastes
// that code into attnpool_cl_device.inc.   attnpool_cl_emu <in.bin> <out.bin>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "simt.h"
#include "scae_b200.h"

namespace scae {
#include "common_device.inc"
#include "attnpool_cl_device.inc"
}  // namespace scae

int main(int argc, char** argv) {
  using namespace scae;
  if (argc != 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int h[5];   // B n D S grid
  if (fread(h, 4, 5, f) != 5) return 2;
  const int B = h[0], n = h[1], D = h[2], S = h[3], grid = h[4], G = D + 1;
  const long groups = (long)B * n;
  if (pool_floats(D, S) > kPoolFloats) return 3;
  std::vector<float> y((size_t)B * S * n * G), g((size_t)groups * D);
  if (fread(y.data(), 4, y.size(), f) != y.size() || fread(g.data(), 4, g.size(), f) != g.size()) return 2;
  fclose(f);
  std::vector<float> out((size_t)groups * D, -7.f), gy(y.size(), -7.f);
  emu_launch(grid, 32 * kPoolWarps, [&] { attnpool_cl_fwd_kernel(y.data(), out.data(), groups, n, D, S); });
  emu_launch(grid, 32 * kPoolWarps, [&] { attnpool_cl_bwd_kernel(y.data(), g.data(), gy.data(), groups, n, D, S); });
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 1;
  fwrite(out.data(), 4, out.size(), o);
  fwrite(gy.data(), 4, gy.size(), o);
  fclose(o);
  return 0;
}
