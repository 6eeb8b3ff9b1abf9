"""Import the UNMODIFIED reference modules from oracle/_ref/reference.zip (built by oracle/build_ref.py).

TEST INFRASTRUCTURE: used by bench.py's reference arm / cpu_baseline leg and by tests; never by the product package.
"""
import hashlib
import importlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
ZIP = os.path.join(HERE, "_ref", "reference.zip")
MANIFEST = os.path.join(HERE, "_ref", "MANIFEST.json")


def available() -> bool:
    return os.path.exists(ZIP) and os.path.exists(MANIFEST)


def verify() -> dict:
    """sha256 of every archive member against the manifest written when it was packed from /root/reference."""
    with open(MANIFEST) as f:
        want = json.load(f)["sha256"]
    with zipfile.ZipFile(ZIP) as z:
        got = {n: hashlib.sha256(z.read(n)).hexdigest() for n in z.namelist()}
    if got != want:
        raise RuntimeError("oracle/_ref/reference.zip does not match its manifest")
    return got


def load(names=("CRN_ELU", "utility")):
    """Returns the requested reference modules (imported from the archive, under the speechbrain / torch_complex shim)."""
    if not available():
        raise FileNotFoundError("oracle/_ref/reference.zip is missing: run `python oracle/build_ref.py` where "
                                "/root/reference is mounted")
    verify()
    import torchaudio
    torchaudio.set_audio_backend = lambda *a, **k: None  # API removed in torchaudio 2.x (reference utility.py:475-476)
    for p in (ZIP, os.path.join(HERE, "shim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    return [importlib.import_module(n) for n in names]
