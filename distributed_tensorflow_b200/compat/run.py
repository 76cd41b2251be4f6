"""``python -m distributed_tensorflow_b200.compat.run script.py [args...]``: run an unmodified TF-1.x script with
``import tensorflow`` resolving to this framework."""
import os
import runpy
import sys

from . import SHIM_DIR


def main():
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m distributed_tensorflow_b200.compat.run script.py [script args]")
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    sys.path.insert(0, SHIM_DIR)
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
