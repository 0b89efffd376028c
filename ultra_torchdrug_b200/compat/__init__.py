"""Import-surface shims for the *unmodified* reference modules.

The reference (`ultra/layer.py`, `ultra/model.py`, `ultra/rel_model.py`) imports
`torchdrug`, `torch_scatter`, `torch_geometric`, `easydict` and the Python-2-era
`collections.Sequence` alias (reference `ultra/model.py:1`, `ultra/rel_model.py:1`).
None of them is installed in this image.  `install()` puts minimal, from-scratch
stand-ins for exactly the attributes those three files touch
on `sys.path`, with `torchdrug.layers.functional.generalized_rspmm` bound to the
B200 operator in `ultra_torchdrug_b200.functional`.

This is plumbing for the drop-in boundary, not a re-implementation of torchdrug.
"""
import collections
import collections.abc
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install():
    """Make `import torchdrug, torch_scatter, torch_geometric, easydict` resolve to the shims."""
    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence  # reference model.py:1 / rel_model.py:1
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    import torchdrug  # noqa: F401  (patches nn.Module.device, see torchdrug/__init__.py)
    return _HERE


def add_reference_to_path(reference_root="/root/reference"):
    """Let `import ultra` find the unmodified reference package (authoring container only)."""
    if not os.path.isdir(os.path.join(reference_root, "ultra")):
        raise FileNotFoundError("reference tree not present at %s" % reference_root)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return reference_root
