// Runs the DEVICE code of torch_scae_b200/csrc/caps_ll2.cu -- the TMA-staged fast path of hot path 2 -- on the CPU:
// tests/emu/simt.h executes the threads, tests/emu/ptx_emu.h stands in for the inline PTX (mbarrier, bulk copies, MUFU),
// the REAL csrc headers (common.cuh, caps_common.cuh) are compiled for the host, and the build script pastes the part
// of caps_ll2.cu above its host-side marker into caps_ll2_device.inc.  The launch sequence of caps2_fwd / caps2_bwd is
// restated below.      caps2_emu <in> <out>
#include <stdio.h>
#include <stdlib.h>

#include "emu_io.h"
#include "simt.h"
#include "ptx_emu.h"
#include "caps_common.cuh"
#undef SCAE_DYNAMIC_SMEM
#define SCAE_DYNAMIC_SMEM(name) float* name = emu_dynamic_smem

namespace scae {
#include "caps_ll2_device.inc"
}  // namespace scae

int main(int argc, char** argv) {
  using namespace scae;
  if (argc != 3) return 1;
  emu_arrays in = emu_read(argv[1]);
  const int* cfg = in["cfg"].as<int>();   // B O V flags threads_fwd threads_bwd grid_bwd
  const int B = cfg[0], O = cfg[1], V = cfg[2], A = 8 * V + 7, P = O * V;
  const unsigned flags = (unsigned)cfg[3];
  const int threads_fwd = cfg[4], threads_bwd = cfg[5], grid_bwd = cfg[6];

  scae_caps_args a;
  a.all_param = in["all_param"].as<float>();
  a.cpr_static = in["cpr_static"].as<float>();
  a.bias_cvr = in["b0"].as<float>();
  a.bias_caps = in["b1"].as<float>();
  a.bias_vote = in["b2"].as<float>();
  a.bias_scale = in["b3"].as<float>();
  a.noise_caps = in["noise_caps"].as<float>();
  a.noise_vote = in["noise_vote"].as<float>();
  a.x = in["x"].as<float>();
  a.presence = in["presence"].as<float>();
  a.dummy_vote = in["dummy_vote"].as<float>();
  a.B = B, a.O = O, a.V = V, a.flags = flags;
  const bool sim = (flags & SCAE_CAPS_SIMILARITY) != 0;

  emu_arrays out;
  const float nanf_ = -777.0f;
  auto f = [&](const char* name, size_t n) {
    out[name] = emu_make<float>(n, nanf_);
    return out[name].as<float>();
  };
  scae_caps_outputs o;
  o.vote = f("vote", (size_t)B * P * 6);
  o.scale = f("scale", (size_t)B * P);
  o.vote_presence = f("vote_presence", (size_t)B * P);
  o.presence_logit_per_caps = f("presence_logit_per_caps", (size_t)B * O);
  o.presence_logit_per_vote = f("presence_logit_per_vote", (size_t)B * P);
  o.caps_presence = f("caps_presence", (size_t)B * O);
  out["caps_presence_arg"] = emu_make<int32_t>((size_t)B * O, -1);
  o.caps_presence_arg = out["caps_presence_arg"].as<int32_t>();
  o.log_prob_per_point = f("log_prob_per_point", (size_t)B * V);
  o.ll_per_example = f("ll_per_example", B);
  o.reg_per_example = f("reg_per_example", B);
  o.vote_presence_binary = f("vote_presence_binary", (size_t)B * P);
  o.winner = f("winner", (size_t)B * V * 6);
  o.winner_presence = f("winner_presence", (size_t)B * V);
  out["winner_idx"] = emu_make<int64_t>((size_t)B * V, -1);
  o.winner_idx = out["winner_idx"].as<int64_t>();
  out["is_from_capsule"] = emu_make<int64_t>((size_t)B * V, -1);
  o.is_from_capsule = out["is_from_capsule"].as<int64_t>();
  o.soft_winner = f("soft_winner", (size_t)B * V * 6);
  o.soft_winner_presence = f("soft_winner_presence", (size_t)B * V);
  o.posterior_mixing_prob = f("posterior_mixing_prob", (size_t)B * P);
  o.mixing_log_prob = f("mixing_log_prob", (size_t)B * (O + 1) * V);
  o.mixing_logit = f("mixing_logit", (size_t)B * (O + 1) * V);

  // ---- forward: one CTA per image (caps2_fwd) ----
  const Caps2FwdLayout LF = caps2_fwd_layout(O, V, a.noise_vote != nullptr);
  if ((size_t)LF.total > kEmuSmemFloats) return 3;
  emu_launch(B, threads_fwd, [&] {
    if (sim) caps2_fwd_kernel<true>(a, o, LF);
    else caps2_fwd_kernel<false>(a, o, LF);
  });

  // ---- backward: persistent CTAs, two stages when they fit (caps2_bwd) ----
  scae_caps_saved sv;
  sv.posterior_mixing_prob = o.posterior_mixing_prob;
  sv.log_prob_per_point = o.log_prob_per_point;
  sv.caps_presence_arg = o.caps_presence_arg;
  sv.winner_idx = o.winner_idx;
  scae_caps_upstream up;
  memset(&up, 0, sizeof(up));
  up.g_ll_per_example = in["g_ll_per_example"].as<float>();
  up.g_reg_per_example = in["g_reg_per_example"].as<float>();
  up.g_posterior_mixing_prob = in["g_posterior_mixing_prob"].as<float>();
  up.g_caps_presence = in["g_caps_presence"].as<float>();
  up.g_vote_presence = in["g_vote_presence"].as<float>();
  up.g_vote = in["g_vote"].as<float>();
  up.g_scale = in["g_scale"].as<float>();
  up.g_presence_logit_per_caps = in["g_presence_logit_per_caps"].as<float>();
  up.g_presence_logit_per_vote = in["g_presence_logit_per_vote"].as<float>();
  up.g_mixing_logit = in["g_mixing_logit"].as<float>();
  const int stages = cfg[7];
  const Caps2BwdLayout LB = caps2_bwd_layout(O, V, a.noise_vote != nullptr, up.g_posterior_mixing_prob != nullptr, stages);
  if ((size_t)LB.total > kEmuSmemFloats) return 3;
  if (V * 8 + O * 4 > kSmallMax * threads_bwd) return 4;
  float* g_all = f("g_all_param", (size_t)B * O * A);
  float* g_presence = a.presence ? f("g_presence", (size_t)B * V) : nullptr;
  std::vector<float> partials((size_t)grid_bwd * O * A, nanf_);
  Caps2BwdOut bo{g_all, g_presence, partials.data()};
  emu_launch(grid_bwd, threads_bwd, [&] {
    if (sim) caps2_bwd_kernel<true>(a, sv, up, bo, LB);
    else caps2_bwd_kernel<false>(a, sv, up, bo, LB);
  });
  float* g_shared = f("g_shared", (size_t)O * A);   // launch_reduce_rows: fixed-order sum over the CTAs' partial rows
  for (int i = 0; i < O * A; ++i) {
    float t = 0.0f;
    for (int c = 0; c < grid_bwd; ++c) t += partials[(size_t)c * O * A + i];
    g_shared[i] = t;
  }
  emu_write(argv[2], out);
  return 0;
}
