"""Copy the UNMODIFIED reference files the benchmark's reference arm drives into baseline/_ref/ (git-ignored, travels to
the GPU box with the gpurun snapshot; nothing under it is ever committed).  Run in the build container, where
/root/reference exists:  python baseline/install_ref.py   (also called by __graft_entry__.build()).

The reference is a pure-Python research repo without packaging metadata (no setup.py / pyproject), so `pip install
--target baseline/_ref /root/reference` has nothing to install; the files SURVEY.md §8c lists for the hot path are
copied byte for byte instead and their SHA-256 digests recorded in baseline/_ref/MANIFEST.json so that "unmodified" can
be checked.  monai (SwinUNETR's dependency) is not in the image, so swin_unetr.py is copied but cannot be imported —
bench.py falls back to the reference-pinned oracle for that workload and says so."""
import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("B200SEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["model/dim3/unet.py", "model/dim3/unet_utils.py", "model/dim3/utils.py", "model/dim3/conv_layers.py",
         "model/dim3/trans_layers.py", "model/dim3/medformer.py", "model/dim3/medformer_utils.py", "model/dim3/unetpp.py",
         "model/dim3/swin_unetr.py", "model/dim3/attention_unet.py", "model/dim3/attention_unet_utils.py",
         "training/losses.py", "training/utils.py", "training/augmentation.py"]


def main():
    if not os.path.isdir(REF):
        print("install_ref: %s not present (GPU box?) — keeping whatever baseline/_ref already holds" % REF)
        return 0
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": REF, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    print("install_ref: %d reference files -> %s" % (len(manifest), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
