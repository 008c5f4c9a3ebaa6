"""B200-native FM / MF training and ranking evaluation (drop-in for the reference's hot path).

Submodules are imported lazily so that host-only helpers (``rfm_b200.synth``) work without
the CUDA library; anything that computes loads ``librfm_b200.so`` and fails loudly without it.
"""
__version__ = "0.1.0"
