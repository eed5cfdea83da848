"""Timeline of one CTA of the tcgen05 GEMM (YNB_TC_TRACE): who waits for whom.
    python tools/gpu_tc_trace.py M K N [mode]
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

CHILD = r'''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import gpu_util as G
from yolo_nano_b200 import _lib
lib = _lib.load()
m, k, n = %d, %d, %d
mode = %d
x = torch.randn(m, k, device=G.DEV); w = torch.randn(n, k, device=G.DEV); b = torch.zeros(n, device=G.DEV)
ld = (n + 3) // 4 * 4
import os
if os.environ.get('YNB_TC_PASS'): ld = 2 * ld
out = torch.zeros(m, ld, device=G.DEV)
for _ in range(2):
    rc = lib.ynb_pwconv_tc(G.ptr(x), k, 0, G.ptr(out), ld, 0, 1, G.ptr(w), G.ptr(b), m, k, n, 1, mode, G.stream())
assert rc == 0
'''


def main():
    m, k, n = (int(v) for v in sys.argv[1:4])
    mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # 1 = 3xTF32, 2 = single-pass TF32
    env = dict(os.environ, YNB_TC_TRACE="1")
    code = CHILD % (str(ROOT), str(ROOT / "tests"), m, k, n, mode)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    lines = res.stderr.splitlines()
    hdrs = [l for l in lines if l.startswith("YNB_TC_TRACE")]
    # keep the events of the LAST (warm) launch
    last = max(i for i, l in enumerate(lines) if l.startswith("YNB_TC_TRACE"))
    ev = [tuple(int(v) for v in l.split()[1:]) for l in lines[last + 1:] if l.startswith("TRACE")]
    print(hdrs[-1])
    if not ev:
        print(res.stderr[-2000:])
        return
    t0 = min(e[3] for e in ev)
    names = {7: "blk(ld,st)", 1: "tma-issued", 2: "split-done", 3: "mma-issued", 4: "epi-wait", 5: "epi-done", 6: "acc-ready"}
    for role, tile, step, clk in sorted(ev, key=lambda e: e[3])[:140]:
        print(f"{clk - t0:8d} cyc  {names.get(role, role):11s} tile {tile:4d} step {step}")
    print(f"... {len(ev)} events, span {max(e[3] for e in ev) - t0} cycles")


if __name__ == "__main__":
    main()
