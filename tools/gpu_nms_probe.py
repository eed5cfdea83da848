"""Times the NMS stage alone on the bench-shaped worst case (reference-init candidates: every
anchor passes the threshold), batch 64.  Run under ncu for per-kernel numbers."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import gpu_util as G  # noqa: E402
from yolo_nano_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    tag = sys.argv[1] if len(sys.argv) > 1 else "refinit"
    g = np.load(ROOT / "tests" / "golden" / f"g3_coco416_{tag}.npz")
    b = 64
    boxes = np.stack([g[f"img{i % 2}.all_bbox"] for i in range(b)])
    scores = np.stack([g[f"img{i % 2}.all_score"] for i in range(b)])
    cls = np.stack([g[f"img{i % 2}.all_cls"] for i in range(b)])
    print("classes populated:", np.bincount(cls[0], minlength=80).tolist())
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        kept, counts, *_ = G.run_nms(lib, boxes, scores, cls, 80, 0.001, 0.5, False,
                                     grid_size=416 if "grid" in sys.argv else 0)
        dt = time.perf_counter() - t0
    print("kept per image", counts[:4], "wall incl. H2D/D2H", dt)


if __name__ == "__main__":
    main()
