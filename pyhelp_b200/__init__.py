"""pyhelp_b200 -- B200-native engine for PyHELP's per-cell daily HELP water balance.

Drop-in for the hot path ``run_help_allcells()`` -> ``HELP3O.run_simulation``
(reference ``pyhelp/processing.py:30-59``, ``pyhelp/HELP3O.FOR``).  See DESIGN.md.
"""
from .processing import (HelpEngine, post_process_output, run_batch, run_help_allcells,  # noqa: F401
                         run_help_singlecell, split_by_cell)
from .inputs import CellBatch, batch_from_cellparams  # noqa: F401

__version__ = "0.1.0"


def patch_pyhelp():
    """Install the GPU path behind an importable reference ``pyhelp`` package:
    ``HelpManager.calc_help_cells`` then calls this module's ``run_help_allcells``
    (reference managers.py:27,440) with no other change."""
    import pyhelp.managers as _m
    import pyhelp.processing as _p
    _m.run_help_allcells = run_help_allcells
    _p.run_help_allcells = run_help_allcells
    _p.run_help_singlecell = run_help_singlecell
