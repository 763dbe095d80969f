#!/usr/bin/env python
"""Audit a GenConViT checkpoint against the state_dict inventory this drop-in (and the reference) expects.

    python tools/audit_checkpoint.py weight/genconvit_ed_inference.pth  --net ed  [--size tiny|large]
    python tools/audit_checkpoint.py weight/genconvit_vae_inference.pth --net vae

The reference loads ``weight/{name}.pth`` -- a raw ``state_dict`` or ``{'state_dict': ...}`` -- with a strict
``load_state_dict`` (reference model/genconvit.py:16-21).  All ConvNeXt / Swin key names are owned by timm==0.6.5, which
is not installable offline, so the names built into this package are restated from timm's published module tree
(SURVEY.md App. A).  This tool is the first-contact check with real weights: it prints every missing, unexpected and
shape-mismatched key with the closest expected name, a per-prefix summary, and exits non-zero if a strict load would
fail.  It only reads tensor shapes (``torch.load(..., mmap=True)`` when the file allows it) and builds the expected
module on the ``meta`` device, so it needs neither a GPU nor memory for a second copy of the weights.
"""
from __future__ import annotations

import argparse
import collections
import difflib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def expected_inventory(net: str, size: str):
    from model.config import load_config
    from model.genconvit_ed import GenConViTED
    from model.genconvit_vae import GenConViTVAE
    cfg = load_config()
    if size != "tiny":
        cfg["model"] = dict(cfg["model"], backbone=f"convnext_{size}", embedder=f"swin_{size}_patch4_window7_224")
    with torch.device("meta"):
        m = (GenConViTED if net == "ed" else GenConViTVAE)(cfg)
    return collections.OrderedDict((k, (tuple(v.shape), v.dtype)) for k, v in m.state_dict().items())


def load_shapes(path: str):
    try:
        ckpt = torch.load(path, map_location="cpu", mmap=True, weights_only=True)
    except Exception:                                        # legacy (non-zip) files cannot be memory-mapped
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
    wrapped = isinstance(ckpt, dict) and "state_dict" in ckpt
    sd = ckpt["state_dict"] if wrapped else ckpt
    extra = sorted(k for k in ckpt if k != "state_dict") if wrapped else []
    return collections.OrderedDict((k, (tuple(v.shape), v.dtype)) for k, v in sd.items() if torch.is_tensor(v)), wrapped, extra


def audit(have, want):
    missing = [k for k in want if k not in have]
    unexpected = [k for k in have if k not in want]
    mismatched = [(k, have[k][0], want[k][0]) for k in want if k in have and have[k][0] != want[k][0]]
    dtype_diff = [(k, have[k][1], want[k][1]) for k in want if k in have and have[k][0] == want[k][0]
                  and have[k][1].is_floating_point != want[k][1].is_floating_point]
    return missing, unexpected, mismatched, dtype_diff


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("checkpoint")
    ap.add_argument("--net", required=True, choices=["ed", "vae"])
    ap.add_argument("--size", default="tiny", choices=["tiny", "large"])
    ap.add_argument("--max-list", type=int, default=40, help="keys listed per category")
    args = ap.parse_args(argv)
    want = expected_inventory(args.net, args.size)
    have, wrapped, extra = load_shapes(args.checkpoint)
    missing, unexpected, mismatched, dtype_diff = audit(have, want)
    print(f"{args.checkpoint}: {'wrapped {state_dict: ...}' if wrapped else 'raw state_dict'}"
          f"{' + ' + ', '.join(extra) if extra else ''}; {len(have)} tensors, expected {len(want)} "
          f"(GenConViT{args.net.upper()}, {args.size})")

    def closest(k, pool):
        m = difflib.get_close_matches(k, pool, n=1, cutoff=0.6)
        return m[0] if m else None

    for title, items in (("missing (expected, not in the file)", missing), ("unexpected (in the file, not expected)", unexpected)):
        print(f"\n{title}: {len(items)}")
        pool = unexpected if items is missing else missing
        for k in items[:args.max_list]:
            c = closest(k, pool)
            shape = (want if items is missing else have)[k][0]
            print(f"  {k}  {shape}" + (f"   ~ {c} {(have if items is missing else want)[c][0]}" if c else ""))
        if len(items) > args.max_list:
            print(f"  ... {len(items) - args.max_list} more")
    print(f"\nshape mismatches: {len(mismatched)}")
    for k, h, w in mismatched[:args.max_list]:
        print(f"  {k}: file {h}, expected {w}")
    if dtype_diff:
        print(f"\ninteger/float type differences: {len(dtype_diff)}")
        for k, h, w in dtype_diff[:args.max_list]:
            print(f"  {k}: file {h}, expected {w}")
    by_prefix = collections.Counter(k.split(".")[0] for k in missing + unexpected + [m[0] for m in mismatched])
    if by_prefix:
        print("\nproblems by top-level module:", dict(by_prefix))
    ok = not (missing or unexpected or mismatched)
    print("\nstrict load_state_dict would " + ("SUCCEED" if ok else "FAIL"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
