"""Recipe that stages the reference's OWN source files for this path under oracle/_ref/ (TEST / BASELINE INFRASTRUCTURE).

    python oracle/vendor_ref.py            # in the build container, where /root/reference exists

The reference (Nano1337/multimodal-clinical) is pure Python, so there is nothing to compile: the "build" of oracle/_ref
is a byte-for-byte copy of the few modules the late-fusion path consists of, with a sha256 manifest.  oracle/_ref/ is
git-ignored (reference sources are never committed) but not gpurun-ignored, so the copy travels to the GPU box, where
/root/reference does not exist.  Only `bench.py --impl reference` / `cpu_baseline` (the LITERAL reference CPU figure
printed beside the vectorised oracle port) and tests may import from it; nothing under multimodal_clinical_b200/ does.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LF_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["existing_algos/QMF.py", "existing_algos/OGM_GE.py", "utils/__init__.py", "utils/EMA.py", "utils/BaseModel.py",
         "cremad/joint_model_qmf.py", "cremad/joint_model_ogm_ge.py", "cremad/ensemble_model_noised.py", "cremad/backbone.py",
         "enrico/joint_model.py", "food101/joint_model_qmf.py"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=True):
    """Copy FILES from the reference tree into oracle/_ref/ and write MANIFEST.json.  Returns False when the reference
    tree is not there (GPU box: the staged copy, if any, is used as it is)."""
    if not os.path.isdir(REF):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = sha256(dst)
    for pkg in ("existing_algos", "cremad", "enrico", "food101"):          # namespace markers (empty; the reference has none for most)
        init = os.path.join(DST, pkg, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest)} reference files under {DST}")
    return True


def verify():
    """True when oracle/_ref holds exactly the files its manifest lists (unmodified)."""
    path = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(path):
        return False
    man = json.load(open(path))["sha256"]
    return all(os.path.exists(os.path.join(DST, rel)) and sha256(os.path.join(DST, rel)) == h for rel, h in man.items())


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
