"""How does tcgen05 kind::tf32 round when it accumulates?  Inputs that are exactly
representable in TF32 make a single pass exact except for the accumulation, so the error
against fp64 isolates the accumulator behaviour.  Prints signed mean / max relative error
vs K for: single-pass TF32 on tf32-exact inputs, 3xTF32 on general fp32 inputs, FFMA."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import gpu_util as G  # noqa: E402
from yolo_nano_b200 import _lib  # noqa: E402


def trunc_tf32(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def main():
    lib = _lib.load()
    torch.manual_seed(0)
    m, n = 4096, 64
    for positive in (False, True):
        print("all-positive operands" if positive else "random-sign operands")
        for k in (32, 64, 128, 232, 464):
            x = torch.randn(m, k, device=G.DEV)
            w = torch.randn(n, k, device=G.DEV) / k ** 0.5
            if positive:
                x, w = x.abs(), w.abs()
            b = torch.zeros(n, device=G.DEV)
            for name, xi, wi, fn, mode in (("tf32 exact-in", trunc_tf32(x), trunc_tf32(w), lib.ynb_pwconv_tc, 2),
                                           ("3xtf32", x, w, lib.ynb_pwconv_tc, 1),
                                           ("ffma", x, w, lib.ynb_pwconv, None)):
                xi, wi = xi.contiguous(), wi.contiguous()
                ref = xi.double() @ wi.double().t()
                out = torch.zeros(m, n, device=G.DEV)
                args = [G.ptr(xi), k, 0, G.ptr(out), n, 0, 1, G.ptr(wi), G.ptr(b), m, k, n, 0]
                if mode is not None:
                    args.append(mode)
                args.append(G.stream())
                assert fn(*args) == 0, lib.ynb_last_error(None)
                torch.cuda.synchronize()
                scale = ref.abs().mean()
                err = (out.double() - ref)
                # signed error in the direction of the result's sign: negative = shrink towards zero
                shrink = (err * torch.sign(ref)).mean() / scale
                print(f"  K={k:3d} {name:14s} mean signed err/scale={shrink:+.3e}  rms/scale={err.pow(2).mean().sqrt() / scale:.3e} "
                      f"max/scale={err.abs().max() / scale:.3e}")


if __name__ == "__main__":
    main()
