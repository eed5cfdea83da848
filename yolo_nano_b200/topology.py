"""Static description of the YOLO-Nano-1.0x detection network.

One table of the 77 convolutions on the path, in the order the engine consumes
them, with the dotted ``state_dict`` prefix each one has in the reference
(`/root/reference/models/yolo_nano.py:32-70`, `backbone/shufflenetv2.py:109-125`).
The Python boundary, the weight packer, the oracle and the C engine
(`csrc/topology.h` builds the same list; `tests/test_abi.py` checks they agree)
all derive from this table, so a layer can never be wired differently in two
places.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

STAGE_REPEATS = (4, 8, 4)                 # backbone/shufflenetv2.py:90
STAGE_CHANNELS = (24, 116, 232, 464)      # backbone/shufflenetv2.py:98 ('1.0x')
NECK_CHANNELS = 96                        # models/yolo_nano.py:40-47
STRIDES = (8, 16, 32)                     # models/yolo_nano.py:23

# activation codes shared with the engine (include/yolonano_b200.h)
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2

# conv kinds shared with the engine
KIND_DENSE3X3, KIND_DW3X3, KIND_PW1X1 = 0, 1, 2


@dataclass(frozen=True)
class ConvSpec:
    name: str            # state_dict prefix of the nn.Conv2d, e.g. "backbone.stage2.0.branch2.3"
    bn: Optional[str]    # state_dict prefix of the BatchNorm that follows (None: head_det_*.4)
    kind: int
    cin: int
    cout: int
    stride: int
    act: int
    conv_bias: bool      # the unfused conv owns a bias parameter

    @property
    def groups(self) -> int:
        return self.cin if self.kind == KIND_DW3X3 else 1

    @property
    def ksize(self) -> int:
        return 1 if self.kind == KIND_PW1X1 else 3

    def weight_shape(self):
        return (self.cout, self.cin // self.groups, self.ksize, self.ksize)


def head_channels(num_classes: int, num_anchors: int = 3) -> int:
    return num_anchors * (1 + num_classes + 4)


def conv_table(num_classes: int, num_anchors: int = 3) -> List[ConvSpec]:
    """All 77 convs in engine order."""
    t: List[ConvSpec] = []

    def add(name, bn, kind, cin, cout, stride, act, bias):
        t.append(ConvSpec(name, bn, kind, cin, cout, stride, act, bias))

    # stem: Conv2d(3,24,3,2,1,bias=False) + BN + ReLU     (shufflenetv2.py:109-113)
    add("backbone.conv1.0", "backbone.conv1.1", KIND_DENSE3X3, 3, STAGE_CHANNELS[0], 2, ACT_RELU, False)
    cin = STAGE_CHANNELS[0]
    for si, (rep, cout) in enumerate(zip(STAGE_REPEATS, STAGE_CHANNELS[1:])):
        h = cout // 2
        for bi in range(rep):
            p = f"backbone.stage{si + 2}.{bi}"
            s = 2 if bi == 0 else 1
            if s == 2:   # branch1 = dw s2 -> BN -> pw -> BN -> ReLU      (shufflenetv2.py:43-49)
                add(f"{p}.branch1.0", f"{p}.branch1.1", KIND_DW3X3, cin, cin, 2, ACT_NONE, False)
                add(f"{p}.branch1.2", f"{p}.branch1.3", KIND_PW1X1, cin, h, 1, ACT_RELU, False)
            b2in = cin if s == 2 else h
            # branch2 = pw -> BN -> ReLU -> dw -> BN -> pw -> BN -> ReLU (shufflenetv2.py:53-63)
            add(f"{p}.branch2.0", f"{p}.branch2.1", KIND_PW1X1, b2in, h, 1, ACT_RELU, False)
            add(f"{p}.branch2.3", f"{p}.branch2.4", KIND_DW3X3, h, h, s, ACT_NONE, False)
            add(f"{p}.branch2.5", f"{p}.branch2.6", KIND_PW1X1, h, h, 1, ACT_RELU, False)
            cin = cout
    n = NECK_CHANNELS
    for i, c in enumerate(STAGE_CHANNELS[1:]):       # lateral 1x1      (yolo_nano.py:40-42)
        add(f"conv1x1_{i}.convs.0", f"conv1x1_{i}.convs.1", KIND_PW1X1, c, n, 1, ACT_LEAKY, True)
    for i in range(4):                               # smooth 3x3       (yolo_nano.py:44-47)
        add(f"smooth_{i}.convs.0", f"smooth_{i}.convs.1", KIND_DENSE3X3, n, n, 1, ACT_LEAKY, True)
    for hd in (1, 2, 3):                             # heads            (yolo_nano.py:50-70)
        p = f"head_det_{hd}"
        add(f"{p}.0.convs.0", f"{p}.0.convs.1", KIND_DW3X3, n, n, 1, ACT_LEAKY, True)
        add(f"{p}.1.convs.0", f"{p}.1.convs.1", KIND_PW1X1, n, n, 1, ACT_LEAKY, True)
        add(f"{p}.2.convs.0", f"{p}.2.convs.1", KIND_DW3X3, n, n, 1, ACT_LEAKY, True)
        add(f"{p}.3.convs.0", f"{p}.3.convs.1", KIND_PW1X1, n, n, 1, ACT_LEAKY, True)
        add(f"{p}.4", None, KIND_PW1X1, n, head_channels(num_classes, num_anchors), 1, ACT_NONE, True)
    assert len(t) == 77
    return t


def num_anchor_boxes(input_size: int, num_anchors: int = 3) -> int:
    """N of SURVEY §8: boxes per image over the three pyramid levels."""
    return num_anchors * sum((input_size // s) ** 2 for s in STRIDES)
