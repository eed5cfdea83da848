#!/usr/bin/env python
"""Timeline of one CTA of the fused depthwise -> pointwise kernel (YNB_DP_TRACE): clock64 stamps per role, tile and
K chunk, printed by the ynb_dwpw_tc test hook.   python tools/gpu_dp_trace.py [batch h w c cout pass(0|1)]"""
import ctypes as C
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["YNB_DP_TRACE"] = "1"
import torch  # noqa: E402
from yolo_nano_b200 import _lib  # noqa: E402

a = [int(v) for v in sys.argv[1:]] or [64, 52, 52, 96, 96, 0, 1]
a = a + [1] * (7 - len(a))
b, h, w, c, cout, with_pass, mode = a          # mode 1 = 3xTF32, 2 = TF32, 3 = bf16
dt = torch.bfloat16 if mode == 3 else torch.float32
al = 8 if mode == 3 else 4
lib = _lib.load()
dev = torch.device("cuda", 0)
p = lambda t: C.c_void_p(t.data_ptr())
x = torch.randn(b, h, w, c, device=dev).to(dt)
dw_w = torch.randn(9, c, device=dev) / 3
dw_b = torch.zeros(c, device=dev)
pw_w = torch.randn(cout, c, device=dev) / c ** 0.5
pw_b = torch.zeros(cout, device=dev)
ld = 2 * cout + 8 if with_pass else cout
out = torch.empty(b, h, w, ld, device=dev, dtype=dt)
x1 = torch.randn(b, h, w, (cout + al - 1) // al * al, device=dev).to(dt) if with_pass else None
for _ in range(2):
    rc = lib.ynb_dwpw_tc(p(x), c, p(dw_w), p(dw_b), 0, p(pw_w), p(pw_b), 1, p(out), ld, p(x1) if with_pass else None,
                         x1.shape[-1] if with_pass else 0, b, h, w, c, cout, mode,
                         C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    assert rc == 0, lib.ynb_last_error(None)
