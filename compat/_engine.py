"""Locate and import the engine package (attention-lvcsr_b200/, registered as attention_lvcsr_b200)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as _graft  # noqa: E402

pkg = _graft.load_package()
