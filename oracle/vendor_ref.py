"""TEST / BENCH INFRASTRUCTURE ONLY — makes the UNMODIFIED reference importable on the GPU box.

`/root/reference` exists only in the build container.  The reference is pure Python, so "building" it means copying the
handful of files its model path imports (verbatim, byte for byte — the manifest records their SHA-256) into the git-ignored
`oracle/_ref/` tree, which travels to the GPU box like the built `.so` files do.  `oracle/ref_shim.py` then imports the
reference from there (same `sys.modules` shims as in the build container), and `bench.py --impl reference` times the
reference's own code (`cpu_baseline.kind = "reference"`) instead of the oracle port.  Nothing under `oracle/_ref/` is ever
committed and nothing in `maed_b200/` imports it.

    python -m oracle.vendor_ref            # called by __graft_entry__.build() when /root/reference is present
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MAED_REFERENCE_SRC", "/root/reference")
# exactly what `import lib.models` pulls in (sys.modules listing after oracle.ref_shim.load_reference()), plus the
# reference's loss for the train-step baseline
FILES = [
    "LICENSE",
    "lib/core/__init__.py", "lib/core/config.py", "lib/core/loss.py",
    "lib/models/__init__.py", "lib/models/ktd.py", "lib/models/maed.py", "lib/models/resnetv2.py", "lib/models/smpl.py",
    "lib/models/spin.py", "lib/models/vision_transformer.py", "lib/models/ops/__init__.py", "lib/models/ops/drop.py",
    "lib/utils/__init__.py", "lib/utils/geometry.py", "lib/utils/utils.py",
]


def vendor(verbose=False):
    """Copies FILES from SRC into oracle/_ref/.  Returns True if the tree is (now) present."""
    if not os.path.isfile(os.path.join(SRC, "lib", "models", "maed.py")):
        return os.path.isfile(os.path.join(DEST, "lib", "models", "maed.py"))
    manifest = {}
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
        if verbose:
            print("vendored", rel)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "note": "verbatim copies of the reference's files (unmodified); git-ignored", "sha256": manifest},
                  f, indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref present:", vendor(verbose=True))
