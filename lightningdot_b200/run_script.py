"""Run one of the reference's scripts UNMODIFIED against this package:

    python -m lightningdot_b200.run_script /path/to/LightningDOT/eval_itm.py  config.json  checkpoint.pt
    torchrun --nproc-per-node 8 -m lightningdot_b200.run_script /path/to/LightningDOT/train_itm.py --config cfg.json

`python eval_itm.py` would put the script's own directory - the reference checkout with ITS dvl/ and uniter_model/ - at
the head of sys.path; this launcher executes the same file with the repository root (the namesake packages dvl,
uniter_model, horovod, apex, GLOBAL_VARIABLES) ahead of it instead, so every import of the script resolves to the B200
implementation.  Nothing in the script is edited or patched.
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    script_dir = os.path.dirname(script)
    sys.path[:] = [ROOT] + [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, script_dir)]
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
