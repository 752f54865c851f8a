#!/usr/bin/env python
"""gpurun_out/parity_<workload>_<mode>.json (written by tests/test_parity_fullsize_gpu.py on the B200 box) -> one markdown
table per workload under profiles/: per-stage error of every math policy against the reference's own forward.
    python tools/parity_table.py gpurun_out profiles/r02_parity_table.md
"""
import glob
import json
import os
import sys


def main(src, dst):
    runs = {}
    for p in sorted(glob.glob(os.path.join(src, "parity_config*_*.json"))):
        r = json.load(open(p))
        runs.setdefault(r["workload"], {})[r["mode"]] = r
    order = ["tf32", "f16", "mixed", "mixed_tf32stereo", "mixed16", "tf32x3", "3xtf32"]
    out = ["# Full-size parity of the product forward against the reference's own forward (round 2, measured on B200)", "",
           "Source: `tests/test_parity_fullsize_gpu.py` (fixtures `tests/golden/golden_config{1,2}.npz` written by the unmodified",
           "reference modules, `oracle/make_golden_full.py`).  Each cell is `max|d|/max|ref|` / `rms(d)/rms(ref)` of the stage's strided",
           "sample, normalised by the reference's full-tensor statistics.  Policies (`stereoscene_b200.ops.MATH_POLICIES`): **tf32** = every",
           "layer plain TF32 on tcgen05; **mixed** (product default) = depth_net + MIE in the compensated TF32x3 mode, the rest plain TF32;",
           "**tf32x3** = every layer compensated on tcgen05; **3xtf32** = every layer compensated on the mma.sync kernels.", ""]
    for wl in sorted(runs):
        modes = [m for m in order if m in runs[wl]]
        out += [f"## {wl}", "", "| stage | " + " | ".join(modes) + " |", "|---|" + "---|" * len(modes)]
        stages = list(runs[wl][modes[0]]["stages"])
        for st in stages:
            out.append(f"| {st} | " + " | ".join(f"{runs[wl][m]['stages'][st]['max']:.2e} / {runs[wl][m]['stages'][st]['rms']:.2e}" for m in modes) + " |")
        out.append("")
        out += ["| against the live CPU oracle (complete tensors) | " + " | ".join(modes) + " |", "|---|" + "---|" * len(modes)]
        for k in ("logits_up_max", "logits_up_rms", "depth_prob_max", "label_agreement", "label_agreement_golden_sample"):
            out.append(f"| {k} | " + " | ".join(f"{runs[wl][m]['live_oracle'][k]:.3e}" if "agree" not in k else f"{runs[wl][m]['live_oracle'][k]:.5f}" for m in modes) + " |")
        out.append("| voxel-index mismatches (of %d points, %d kept) | " % (runs[wl][modes[0]]["points"], runs[wl][modes[0]]["kept_points"]) +
                   " | ".join(str(runs[wl][m]["index_mismatch"]) for m in modes) + " |")
        out.append("| geometry bit-exact vs oracle | " + " | ".join(str(runs[wl][m]["live_oracle"]["geom_bit_exact"]) for m in modes) + " |")
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
