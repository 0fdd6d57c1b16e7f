"""Recipe for `baseline/_ref`: the UNMODIFIED reference model code, placed where the GPU box can import it.

The reference (autonomousvision/gta @ 1dcbd37) is a plain Python tree without setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` has nothing to build (it fails with "neither 'setup.py' nor
'pyproject.toml' found"); this script is the equivalent install: it copies, byte for byte, the files the GTA path and its
callers need —

    source/__init__.py  source/layers.py  source/encoder.py  source/decoder.py  source/models_nvs.py
    source/utils/{__init__,gta,wigner_d,common,nerf,gbt,frustum_posemb}.py
    runs/{clevrtr,msn}/GTA/*/config.yaml
    J_dense.pt   (only J[0..2]: the shipped configs use Wigner-D degrees <= 2; the full file is 36.7 MB)

— into baseline/_ref, which is git-ignored (never part of this repository's history) but travels to the GPU box with
the gpurun snapshot.  Nothing here is imported by the product package; consumers are bench.py (`--impl reference`,
the informational eager-on-GPU leg, `--workload train_step`) and the tests that run the reference's own `Attention`
module against the drop-in.  source/data (TensorFlow / sunds datasets), trainer.py (matplotlib) and the CLI are not
copied: synthetic batches replace them.

    python baseline/install_ref.py [--ref /root/reference]
"""
from __future__ import annotations

import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["source/__init__.py", "source/layers.py", "source/encoder.py", "source/decoder.py", "source/models_nvs.py",
         "source/utils/__init__.py", "source/utils/gta.py", "source/utils/wigner_d.py", "source/utils/common.py",
         "source/utils/nerf.py", "source/utils/gbt.py", "source/utils/frustum_posemb.py"]


def install(ref: str = "/root/reference", dest: str = DEST) -> str:
    if not os.path.isfile(os.path.join(ref, "source", "utils", "gta.py")):
        raise FileNotFoundError("reference tree not found at %s" % ref)
    files = list(FILES) + [os.path.relpath(p, ref) for p in glob.glob(os.path.join(ref, "runs", "*", "GTA", "*", "config.yaml"))]
    for rel in files:
        src, dst = os.path.join(ref, rel), os.path.join(dest, rel)
        if not os.path.exists(src):        # source/__init__.py etc. may be absent (namespace package)
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    jd = os.path.join(dest, "J_dense.pt")
    if not os.path.exists(jd):
        import torch
        J = torch.load(os.path.join(ref, "J_dense.pt"))
        torch.save([j.clone() for j in J[:3]], jd)
    with open(os.path.join(dest, "INSTALLED_FROM"), "w") as f:
        f.write("%s (files copied unmodified by baseline/install_ref.py; J_dense.pt trimmed to degrees 0..2)\n" % ref)
    return dest


if __name__ == "__main__":
    ref = sys.argv[sys.argv.index("--ref") + 1] if "--ref" in sys.argv else "/root/reference"
    print(install(ref))
