"""Micro-benchmark of the tcgen05 pointwise kernel on the network's shapes (bench batch 64).
NOTE: ynb_pwconv_tc is the unit-test hook (it re-packs weights and synchronises per call), so
the kernel is timed with the device trace clock, not wall time: YNB_TC_TRACE gives the span
of CTA 0 in SM cycles.  Prints cycles per tile and the implied per-SM bandwidth."""
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SHAPES = [(173056, 60, 58, "stage2 pw1"), (692224, 24, 58, "stage2.0 b2.0"), (43264, 116, 116, "stage3 pw1"),
          (10816, 232, 232, "stage4 pw1"), (173056, 96, 96, "head1 pw"), (173056, 96, 255, "head1 final"),
          (173056, 120, 96, "lateral c3")]


def main():
    for m, k, n, name in SHAPES:
        res = subprocess.run([sys.executable, str(ROOT / "tools" / "gpu_tc_trace.py"), str(m), str(k), str(n)],
                             capture_output=True, text=True, env=dict(os.environ))
        hdr = [l for l in res.stdout.splitlines() if l.startswith("YNB_TC_TRACE")]
        span = re.search(r"span (\d+) cycles", res.stdout)
        dones = [int(l.split()[0]) for l in res.stdout.splitlines() if "epi-done" in l]
        tiles_cta = len(dones)
        per_tile = (dones[-1] - dones[0]) / max(1, tiles_cta - 1) if tiles_cta > 1 else float("nan")
        byts = 128 * (k + n) * 4
        print(f"{name:14s} M={m:7d} K={k:3d} N={n:3d} {hdr[-1][13:] if hdr else '?':70s} span={span.group(1) if span else '?':>7s} cyc "
              f"tiles/CTA(shown)={tiles_cta:2d} cyc/tile={per_tile:8.0f}  -> {byts / per_tile * 1.9:6.1f} GB/s/SM" if tiles_cta > 1 else
              f"{name:14s} {hdr[-1] if hdr else res.stderr[-300:]} span={span.group(1) if span else '?'}")


if __name__ == "__main__":
    main()
