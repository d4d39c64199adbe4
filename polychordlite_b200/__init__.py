"""polychordlite_b200 -- B200-native nested sampling behind PolyChordLite's API.

The compute path is the CUDA library polychordlite_b200/lib/libchord.so (built by
polychordlite_b200._build); this package is the thin host-side mirror of the reference's
pypolychord interface plus the ctypes binding of the C ABI (include/polychord_b200.h).
"""
__version__ = "0.1.0"
