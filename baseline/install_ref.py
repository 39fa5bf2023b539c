"""Install the reference for the baseline arms: copy the files of the hot path UNMODIFIED from the read-only mount into
``baseline/_ref/reference/`` (git-ignored; it travels to the GPU box with the snapshot, like a pip-installed baseline
would).  The reference has no setup.py / build system, so "install" is a copy of its source files:

    src/__init__.py, src/models.py, src/correlation.py, LICENSE, the demo pair + its shipped output .flo, and the pretrained
    state_dict files if the mount has them (it does not: .MISSING_LARGE_BLOBS)

(``inference.py`` / ``run.py`` import cv2, imutils, torchvision and the plotting stack at module level and are not needed:
``estimate`` is 12 arithmetic lines around ``net(img1, img2)``.)

    python baseline/install_ref.py          (run by __graft_entry__.build(); a no-op without /root/reference)

TEST / BASELINE INFRASTRUCTURE: the product package never reads baseline/.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "reference")
SRC = os.environ.get("PIVLFN_REFERENCE", "/root/reference")
FILES = ["src/__init__.py", "src/models.py", "src/correlation.py", "LICENSE",
         # the reference's one golden artefact (needs the pretrained blob, absent from the mount: tests/test_demo_blob.py lights up
         # when models/pretrain_torch/PIV-LiteFlowNet-en.paramOnly appears) and, if ever present, the weight blobs themselves
         "images/demo/DNS_turbulence_img1.tif", "images/demo/DNS_turbulence_img2.tif", "images/demo/DNS_turbulence_out.flo",
         "images/demo/DNS_turbulence_flow.flo",
         "models/pretrain_torch/PIV-LiteFlowNet-en.paramOnly", "models/pretrain_torch/Hui-LiteFlowNet.paramOnly"]


def install() -> str:
    if not os.path.isfile(os.path.join(SRC, "src", "models.py")):
        return DST                      # GPU box: use what was shipped
    manifest = {}
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        data = open(s, "rb").read()
        manifest[f] = hashlib.sha256(data).hexdigest()
        if not os.path.isfile(d) or open(d, "rb").read() != data:
            shutil.copyfile(s, d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest, "note": "byte-identical copies of the reference's files"}, fh, indent=1)
    return DST


def installed() -> bool:
    return os.path.isfile(os.path.join(DST, "src", "models.py")) and os.path.isfile(os.path.join(DST, "src", "correlation.py"))


if __name__ == "__main__":
    print(install(), "installed" if installed() else "NOT installed")
