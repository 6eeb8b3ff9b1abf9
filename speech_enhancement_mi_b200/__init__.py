"""B200-native streaming speech enhancement: the CRN_ELU realtime_process hot path behind the reference's model API.

Modules mirror the reference's file names so that `from speech_enhancement_mi_b200 import CRN_ELU` (or putting this
directory first on sys.path) is a drop-in for the reference's `import CRN_ELU`.
"""
__all__ = ["CRN_ELU", "distillation_crn", "utility"]
