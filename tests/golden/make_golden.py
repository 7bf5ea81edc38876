#!/usr/bin/env python
"""Generates tests/golden/golden.json from the reference's own compiled C path (oracle/_ref, built from /root/reference by
oracle/build_ref.sh).  Run here, where the reference exists; the digests are committed and travel to the GPU box."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import binding  # noqa: E402
import cases  # noqa: E402


def main():
    ref = binding.reference()
    if ref is None:
        raise SystemExit("oracle/_ref is not built (needs /root/reference)")
    b = cases.CpuBackend(ref)
    out = {name: fn(b) for name, fn in cases.cases()}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"generator": "reference C path, HEVCASM_C_REF|HEVCASM_C_OPT (oracle/_ref/libhevcasm_cref.so)", "sha256": out}, f, indent=1, sort_keys=True)
    print(f"wrote {len(out)} digests")


if __name__ == "__main__":
    main()
