"""Tie-break sensitivity of DistributeOctTree (hard part 1): the reference sorts pair<int, ExtractorNode*>
(S/ORBextractor.cc:694-698), so nodes with equal point counts are ordered by heap address and its output depends on the
allocator.  Parity everywhere else in this repository is against the canonical "address order = creation order" (the
reference under a monotonic bump allocator, oracle/_ref/libref_orb.so).  This script runs the SAME reference build under
glibc's own malloc (oracle/_ref/libref_orb_malloc.so) on the bench frames and reports how far the two differ: what an
integrator who links the real reference will see against the B200 path.

    python tools/tiebreak_report.py [frames] > profiles/r2_tiebreak.json        (build container only: needs oracle/_ref)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_lib as R  # noqa: E402
from weiner_slamit_v2_b200.frames import synthetic_frame  # noqa: E402


def compare(frames, params=(1000, 1.2, 8, 20, 7)):
    canon, glibc = R.RefExtractor(*params), R.RefExtractor(*params, default_malloc=True)
    same_set = same_order = 0
    differing, total, per_frame = 0, 0, []
    for img in frames:
        ka, da = canon(img)
        kb, db = glibc(img)
        ra = [ka[i].tobytes() + da[i].tobytes() for i in range(len(ka))]
        rb = [kb[i].tobytes() + db[i].tobytes() for i in range(len(kb))]
        sa, sb = set(ra), set(rb)
        only = len(sa - sb) + len(sb - sa)                      # keypoints (record + descriptor) present on one side only
        same_set += only == 0
        same_order += ra == rb
        differing += only
        total += len(ra)
        per_frame.append({"keypoints": len(ra), "keypoints_glibc": len(rb), "not_in_both": only, "same_order": ra == rb})
    n = len(frames)
    return {"frames": n, "keypoints_canonical_total": total, "frames_identical_set": same_set, "frames_identical_order": same_order,
            "fraction_identical_set": same_set / n, "fraction_identical_order": same_order / n,
            "keypoints_not_in_both_total": differing, "fraction_of_keypoints_not_in_both": differing / (2.0 * total) if total else 0.0,
            "per_frame": per_frame}


def self_consistency(frames, params=(1000, 1.2, 8, 20, 7)):
    """The glibc-malloc reference against ITSELF: the same frames through one extractor in forward and through another in
    reverse order, i.e. the same image met in two different heap states."""
    fwd, rev = R.RefExtractor(*params, default_malloc=True), R.RefExtractor(*params, default_malloc=True)
    a = [fwd(img) for img in frames]
    b = [rev(img) for img in reversed(frames)][::-1]
    differ = sum(1 for (ka, da), (kb, db) in zip(a, b) if ka.tobytes() != kb.tobytes() or da.tobytes() != db.tobytes())
    return {"frames": len(frames), "frames_whose_output_depends_on_heap_history": differ}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    frames = [synthetic_frame(i) for i in range(n)]              # rank 0's bench frames (bench.py:_frames)
    vga = compare(frames)
    hd = compare([synthetic_frame(1000 * 100000 + i, 1280, 720) for i in range(8)], (2000, 1.2, 8, 20, 7))
    out = {"what": "reference ORBextractor.cc under glibc malloc against the same build under a monotonic bump allocator (the canonical "
                   "order the B200 path reproduces bit for bit)",
           "reference_lines": "S/ORBextractor.cc:694-698 (sort of pair<int, ExtractorNode*>), :743-744",
           "vga_640x480_n1000": vga, "hd_1280x720_n2000": hd,
           "glibc_reference_against_itself": self_consistency(frames),
           "reading": "a keypoint 'not in both' is one whose quadtree leaf was split in one run and not in the other because two nodes "
                      "with equal point counts swapped places in the largest-first order; pyramid, FAST scores, angles and descriptors "
                      "of the keypoints both runs keep are identical"}
    json.dump(out, sys.stdout, indent=1)
    sys.stdout.write("\n")


if __name__ == "__main__":
    main()
