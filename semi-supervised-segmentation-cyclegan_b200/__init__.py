"""B200-native hot path of arnab39/Semi-supervised-segmentation-cycleGAN.

Import name: ``sscg_b200`` (the directory name required by the build contract contains hyphens; the
root-level ``sscg_b200.py`` shim registers this package under the importable name).

Public surface mirrors the reference's network boundary (reference arch/__init__.py:1-3):
``define_Gen``, ``define_Dis``, ``set_grad``.
"""
__version__ = "0.1.0"
