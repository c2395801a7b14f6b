#!/usr/bin/env python
"""Run the reference's own entry point with the B200-native `models` package in place of its own.

    cd /path/to/sr-pytorch-lightning
    python /path/to/this/repo/scripts/run_reference_main.py fit --model RCAN --config configs/train_default_sr.yml ...

`python main.py` puts the script's directory FIRST on sys.path, ahead of PYTHONPATH, so exporting PYTHONPATH alone does
not shadow the reference's `models` package.  This launcher puts sr-pytorch-lightning_b200/ first and then executes
main.py from the current directory (runpy.run_path does not touch sys.path): `import models` inside main.py
(main.py:7,87-93) resolves to the B200 classes, while `srdata`, `losses`, `utils` still come from the reference."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), "sr-pytorch-lightning_b200")


def main():
    ref = os.getcwd()
    script = os.path.join(ref, "main.py")
    if not os.path.isfile(script):
        raise SystemExit("run this from the root of a sr-pytorch-lightning checkout (main.py not found in the current directory)")
    sys.path[:] = [PKG, ref] + [p for p in sys.path if p not in (PKG, ref, HERE, "")]
    sys.argv = [script] + sys.argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
