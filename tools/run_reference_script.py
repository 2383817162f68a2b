#!/usr/bin/env python
"""Runs one of the reference's OWN entry scripts, unmodified, from baseline/_ref:

    python tools/run_reference_script.py [--native] [--synthetic-dataset] pred_test.py --model vmn_gca --load ckpt ...
    python tools/run_reference_script.py --native --synthetic-dataset train_ddp.py --cfg x.yaml --local_rank 0

--native calls tcvom_b200.install() first, so `models.VMN.get_VMN_models('vmn_gca')`,
`models.model.EvalModel` and `FullModel_VMD` resolve to the B200-native implementation; without it the script runs the
reference's PyTorch modules (the comparison arm).  The script then runs under runpy exactly as `python script.py ...`."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    argv = sys.argv[1:]
    native = synthetic = False
    while argv and argv[0].startswith("--"):
        flag = argv.pop(0)
        if flag == "--native":
            native = True
        elif flag == "--synthetic-dataset":
            synthetic = True
        else:
            raise SystemExit(f"unknown runner flag {flag}")
    if not argv:
        raise SystemExit(__doc__)
    from baseline import ref_env
    ref = ref_env.activate(cpu=False, synthetic_dataset=synthetic)
    if native:
        import tcvom_b200
        tcvom_b200.install()
    script = os.path.join(ref, argv[0])
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
