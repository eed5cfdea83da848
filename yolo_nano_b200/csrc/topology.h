// The 77 convolutions of the YOLO-Nano-1.0x path in engine order (C++ twin of
// yolo_nano_b200/topology.py; tests/test_abi.py checks that both agree).
// Reference: backbone/shufflenetv2.py:109-125, models/yolo_nano.py:40-70.
#pragma once
#include <string>
#include <vector>

#include "yolonano_b200.h"

namespace ynb {

enum ConvKind { kDense3x3 = 0, kDw3x3 = 1, kPw1x1 = 2 };

struct ConvSpec {
  std::string name;
  ConvKind kind;
  int cin, cout, stride, act;
};

inline const int* stage_channels() {
  static const int c[4] = {24, 116, 232, 464};
  return c;
}
inline const int* stage_repeats() {
  static const int r[3] = {4, 8, 4};
  return r;
}
constexpr int kNeckC = 96;

inline std::vector<ConvSpec> conv_table(int num_classes, int num_anchors) {
  std::vector<ConvSpec> t;
  auto add = [&](std::string n, ConvKind k, int ci, int co, int s, int act) {
    t.push_back(ConvSpec{std::move(n), k, ci, co, s, act});
  };
  add("backbone.conv1.0", kDense3x3, 3, 24, 2, YNB_ACT_RELU);
  int cin = 24;
  for (int si = 0; si < 3; ++si) {
    int cout = stage_channels()[si + 1], h = cout / 2;
    for (int bi = 0; bi < stage_repeats()[si]; ++bi) {
      std::string p = "backbone.stage" + std::to_string(si + 2) + "." + std::to_string(bi);
      int s = bi == 0 ? 2 : 1;
      if (s == 2) {
        add(p + ".branch1.0", kDw3x3, cin, cin, 2, YNB_ACT_NONE);
        add(p + ".branch1.2", kPw1x1, cin, h, 1, YNB_ACT_RELU);
      }
      add(p + ".branch2.0", kPw1x1, s == 2 ? cin : h, h, 1, YNB_ACT_RELU);
      add(p + ".branch2.3", kDw3x3, h, h, s, YNB_ACT_NONE);
      add(p + ".branch2.5", kPw1x1, h, h, 1, YNB_ACT_RELU);
      cin = cout;
    }
  }
  for (int i = 0; i < 3; ++i)
    add("conv1x1_" + std::to_string(i) + ".convs.0", kPw1x1, stage_channels()[i + 1], kNeckC, 1, YNB_ACT_LEAKY);
  for (int i = 0; i < 4; ++i)
    add("smooth_" + std::to_string(i) + ".convs.0", kDense3x3, kNeckC, kNeckC, 1, YNB_ACT_LEAKY);
  for (int hd = 1; hd <= 3; ++hd) {
    std::string p = "head_det_" + std::to_string(hd);
    add(p + ".0.convs.0", kDw3x3, kNeckC, kNeckC, 1, YNB_ACT_LEAKY);
    add(p + ".1.convs.0", kPw1x1, kNeckC, kNeckC, 1, YNB_ACT_LEAKY);
    add(p + ".2.convs.0", kDw3x3, kNeckC, kNeckC, 1, YNB_ACT_LEAKY);
    add(p + ".3.convs.0", kPw1x1, kNeckC, kNeckC, 1, YNB_ACT_LEAKY);
    add(p + ".4", kPw1x1, kNeckC, num_anchors * (1 + num_classes + 4), 1, YNB_ACT_NONE);
  }
  return t;
}

}  // namespace ynb
