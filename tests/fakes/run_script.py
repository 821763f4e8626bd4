"""Runs an UNMODIFIED script (e.g. one of the reference's) with the Gallery test double installed:

    python tests/fakes/run_script.py /path/to/script.py

Used by the CPU conformance tests only; on a B200 the scripts run directly against librbod.so.
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))

from fakes import fake_gallery  # noqa: E402

fake_gallery.install()
script = sys.argv[1]
sys.argv = [script] + sys.argv[2:]
runpy.run_path(script, run_name="__main__")
