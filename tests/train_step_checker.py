"""In-chain checker for `yolo_nano_b200.train_step.TrainStep` (test infrastructure; also used by
tools/gpu_train_step_debug.py)."""
import torch
import torch.nn.functional as F

from yolo_nano_b200.train_step import TrainStep

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


class CheckedTrainStep(TrainStep):
    """TrainStep whose every backward closure is compared, inside the running chain, with torch.autograd of the same
    op on the same inputs and incoming gradient (float32, TF32 off).  torch is the CHECKER here, not the product."""

    def __init__(self, model):
        super().__init__(model)
        self.records = []
        self.check_flags = True

    def _wrap(self, tape, name, x, y, ref_fn, pnames):
        """ref_fn(x_req) -> y_ref, [param tensors requiring grad] ; compares dx and the parameter gradients."""
        inner = tape.ops.pop()

        def bw():
            dy = tape.grad[id(y)].clone()
            captured = []
            orig = tape.add_grad

            def cap(t, g):
                captured.append((t, g.clone()))
                orig(t, g)
            tape.add_grad = cap
            inner()
            tape.add_grad = orig
            with torch.enable_grad():
                xr = x.detach().clone().requires_grad_(True)
                yr, ps = ref_fn(xr)
                gs = torch.autograd.grad(yr, [xr] + ps, dy)
            fwd = rel(y, yr.detach())
            rec = {"op": name, "fwd": fwd, "dx": rel(captured[0][1], gs[0]), "params": {}}
            scale = max([float(g.abs().max()) for g in gs[1:]] + [1e-30])
            for pn, g in zip(pnames, gs[1:]):
                # error relative to the largest gradient of the op's parameters (a conv bias in front of a
                # BatchNorm has gradient exactly 0 in real arithmetic, rounding noise in float32)
                rec["params"][pn] = float((tape.pgrad[pn] - g).abs().max()) / scale
            self.records.append(rec)
        tape.ops.append(bw)

    def pw(self, tape, x, conv, bias):
        y = super().pw(tape, x, conv, bias)
        w = self.sd[conv + ".weight"].clone()
        b = self.sd[conv + ".bias"].clone() if bias else None
        n, k = w.shape[0], w.shape[1]

        def ref(xr):
            wr = w.requires_grad_(True)
            ps = [wr]
            br = None
            if bias:
                br = b.requires_grad_(True)
                ps.append(br)
            o = F.conv2d(xr[..., :k].permute(0, 3, 1, 2), wr, br).permute(0, 2, 3, 1)
            o = F.pad(o, (0, y.shape[-1] - n))
            return o, ps
        self._wrap(tape, "pw " + conv, x, y, ref, [conv + ".weight"] + ([conv + ".bias"] if bias else []))
        return y

    def dw(self, tape, x, conv, stride, bias):
        y = super().dw(tape, x, conv, stride, bias)
        w = self.sd[conv + ".weight"].clone()
        b = self.sd[conv + ".bias"].clone() if bias else None
        c = w.shape[0]

        def ref(xr):
            wr = w.requires_grad_(True)
            ps = [wr]
            br = None
            if bias:
                br = b.requires_grad_(True)
                ps.append(br)
            o = F.conv2d(xr[..., :c].permute(0, 3, 1, 2), wr, br, stride=stride, padding=1, groups=c).permute(0, 2, 3, 1)
            return F.pad(o, (0, y.shape[-1] - c)), ps
        self._wrap(tape, "dw " + conv, x, y, ref, [conv + ".weight"] + ([conv + ".bias"] if bias else []))
        return y

    def conv3(self, tape, x, conv):
        y = super().conv3(tape, x, conv)
        w, b = self.sd[conv + ".weight"].clone(), self.sd[conv + ".bias"].clone()

        def ref(xr):
            wr, br = w.requires_grad_(True), b.requires_grad_(True)
            return F.conv2d(xr.permute(0, 3, 1, 2), wr, br, padding=1).permute(0, 2, 3, 1), [wr, br]
        self._wrap(tape, "conv3 " + conv, x, y, ref, [conv + ".weight", conv + ".bias"])
        return y

    def bn(self, tape, x, bn_name, act):
        y = super().bn(tape, x, bn_name, act)
        g, b = self.sd[bn_name + ".weight"].clone(), self.sd[bn_name + ".bias"].clone()
        c = g.shape[0]

        def ref(xr):
            gr, br = g.requires_grad_(True), b.requires_grad_(True)
            o = F.batch_norm(xr[..., :c].permute(0, 3, 1, 2), None, None, gr, br, True, 0.1, 1e-5)
            o = {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, 0.1)}[act](o).permute(0, 2, 3, 1)
            return F.pad(o, (0, y.shape[-1] - c)), [gr, br]
        self._wrap(tape, "bn%d %s" % (act, bn_name), x, y, ref, [bn_name + ".weight", bn_name + ".bias"])
        return y

    def merge(self, tape, a, a2, mode):
        y = super().merge(tape, a, a2, mode)
        inner = tape.ops.pop()

        def bw():
            dy = tape.grad[id(y)].clone()
            captured = {}
            orig = tape.add_grad

            def cap(t, g):
                captured[id(t)] = g.clone()
                orig(t, g)
            tape.add_grad = cap
            inner()
            tape.add_grad = orig
            with torch.enable_grad():
                ar, a2r = a.detach().clone().requires_grad_(True), a2.detach().clone().requires_grad_(True)
                yr = ar + F.interpolate(a2r.permute(0, 3, 1, 2), scale_factor=2.0 if mode == 1 else 0.5).permute(0, 2, 3, 1)
                ga, ga2 = torch.autograd.grad(yr, [ar, a2r], dy)
            self.records.append({"op": "merge mode %d %s" % (mode, tuple(a.shape)), "fwd": rel(y, yr.detach()),
                                 "dx": max(rel(captured[id(a)], ga), rel(captured[id(a2)], ga2)), "params": {}})
        tape.ops.append(bw)
        return y
