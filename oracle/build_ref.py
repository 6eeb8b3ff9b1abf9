"""Recipe for oracle/_ref/: pack the UNMODIFIED reference sources of the hot path so that they travel to the GPU box.

TEST INFRASTRUCTURE (the checker / the reference arm of bench.py), never imported by the product package.

    python oracle/build_ref.py          # needs /root/reference (this container); a no-op elsewhere

The reference is pure Python (no native code to compile), so "building" it means packing the four files the path
lives in -- CRN_ELU.py, distillation_crn.py, fullsubnet.py, utility.py -- byte for byte into
oracle/_ref/reference.zip (git-ignored: the sources never enter this repository's history; Python imports straight
from the archive) together with MANIFEST.json (sha256 of every member, so that "unmodified" can be checked).  The
reference's un-vendored dependency speechbrain (STFT / ISTFT wrappers over torch.stft / torch.istft) and the unused
torch_complex import are served by oracle/shim/ (SURVEY.md appendix A).  `oracle/ref_loader.py` imports the archive.
"""
import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SE_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["CRN_ELU.py", "distillation_crn.py", "fullsubnet.py", "utility.py"]


def build(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"{REF} is absent: keeping whatever oracle/_ref/ already holds")
        return os.path.exists(os.path.join(OUT, "reference.zip"))
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    tmp = os.path.join(OUT, "reference.zip.tmp")
    with zipfile.ZipFile(tmp, "w", compression=zipfile.ZIP_DEFLATED) as z:
        for name in FILES:
            with open(os.path.join(REF, name), "rb") as f:
                data = f.read()
            manifest[name] = hashlib.sha256(data).hexdigest()
            # fixed timestamp: the archive is reproducible byte for byte
            z.writestr(zipfile.ZipInfo(name, date_time=(2020, 1, 1, 0, 0, 0)), data)
    os.replace(tmp, os.path.join(OUT, "reference.zip"))
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    if verbose:
        print("oracle/_ref/reference.zip:", manifest)
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
