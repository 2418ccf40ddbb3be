// Runs the DEVICE code of torch_scae_b200/csrc/loss_head.cu on the CPU (tests/emu/simt.h) -- the build script pastes
// that code into loss_head_device.inc -- with the launch sequence of scae_loss_head_fwd / _bwd restated below.
//   loss_head_emu <in.bin> <out.bin>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "simt.h"
#include "scae_b200.h"

namespace scae {
#include "common_device.inc"
#include "loss_head_device.inc"
}  // namespace scae

template <class T>
static std::vector<T> read_vec(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) {
    fprintf(stderr, "short read\n");
    exit(2);
  }
  return v;
}

int main(int argc, char** argv) {
  using namespace scae;
  if (argc != 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  // header: B O V K has_label sparsity prior_type posterior_type grid | 7 floats (weights x4, constants x3) | g_total
  const std::vector<int> h = read_vec<int>(f, 9);
  const std::vector<float> c = read_vec<float>(f, 8);
  const int B = h[0], O = h[1], V = h[2], K = h[3], has_label = h[4], grid = h[8];
  std::vector<float> cp = read_vec<float>(f, (size_t)B * O), post = read_vec<float>(f, (size_t)B * O * V);
  std::vector<long long> label = read_vec<long long>(f, has_label ? B : 0);
  std::vector<float> w = read_vec<float>(f, has_label ? (size_t)K * O : 0), bias = read_vec<float>(f, has_label ? K : 0);
  fclose(f);
  scae_loss_head_args a;
  a.caps_presence = cp.data();
  a.posterior = post.data();
  a.label = has_label ? label.data() : nullptr;
  a.cls_weight = has_label ? w.data() : nullptr;
  a.cls_bias = has_label ? bias.data() : nullptr;
  a.B = B, a.O = O, a.V = V, a.K = K;
  a.sparsity = h[5], a.prior_type = h[6], a.posterior_type = h[7];
  a.prior_within_weight = c[0], a.prior_between_weight = c[1];
  a.posterior_within_weight = c[2], a.posterior_between_weight = c[3];
  a.prior_within_constant = c[4], a.posterior_within_constant = c[5], a.between_constant = c[6];
  const float g_total = c[7];
  const int n_cls = has_label ? K * O + K : 0;
  const int vec = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(post.data()) & 15u) == 0);

  std::vector<float> terms(8, -1.f), stats(128, -1.f), probs((size_t)2 * B * (has_label ? K : 0), -1.f);
  std::vector<float> partials((size_t)grid * (n_cls > kHeadStat ? n_cls : kHeadStat), -1.f);
  emu_launch(grid, 32 * kHeadWarps,
             [&] { loss_head_fwd_kernel(a, has_label ? probs.data() : nullptr, partials.data(), vec); });
  emu_launch(1, 256, [&] { loss_head_finalize_kernel(a, partials.data(), grid, terms.data(), stats.data()); });

  std::vector<float> g_cp((size_t)B * O, -1.f), g_post((size_t)B * O * V, -1.f), g_cls(n_cls, 0.f);
  std::vector<float> bpart((size_t)grid * (n_cls > kHeadStat ? n_cls : kHeadStat), -1.f);
  emu_launch(grid, 32 * kHeadWarps, [&] {
    loss_head_bwd_kernel(a, stats.data(), &g_total, a.sparsity ? g_cp.data() : nullptr,
                         a.sparsity ? g_post.data() : nullptr, bpart.data(), vec);
  });
  for (int i = 0; i < n_cls; ++i)   // launch_reduce_rows: fixed-order sum over the CTAs' partial rows
    for (int p = 0; p < grid; ++p) g_cls[i] += bpart[(size_t)p * n_cls + i];

  FILE* o = fopen(argv[2], "wb");
  if (!o) return 1;
  fwrite(terms.data(), 4, 8, o);
  fwrite(probs.data(), 4, probs.size(), o);
  fwrite(g_cp.data(), 4, g_cp.size(), o);
  fwrite(g_post.data(), 4, g_post.size(), o);
  fwrite(g_cls.data(), 4, g_cls.size(), o);
  fclose(o);
  return 0;
}
