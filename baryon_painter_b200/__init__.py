"""baryon_painter_b200 -- B200-native paint path of tilmantroester/baryon_painter.

Drop-in surface (same names as the reference package):
    baryon_painter_b200.painter.CVAEPainter / CGANPainter   -> .paint(input, z=, transform=, inverse_transform=)
    baryon_painter_b200.process_SLICS.process_SLICS / create_y_map / generate_tiling / get_tile / make_weight_map
"""

__version__ = "0.1.0"


def pinned_empty(shape, dtype="float32"):
    """Page-locked numpy array for zero-staging host transfers (see ``_lib.pinned_empty``)."""
    from ._lib import pinned_empty as _pe
    return _pe(shape, dtype)
