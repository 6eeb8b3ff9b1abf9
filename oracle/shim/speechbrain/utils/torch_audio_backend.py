"""Import-time stub for the un-vendored speechbrain dependency (reference utility.py:473-476)."""


def get_torchaudio_backend():
    return "soundfile"
