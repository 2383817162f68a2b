#!/usr/bin/env python
"""The reference's calc_metric.py (baseline/_ref, unmodified) with its per-frame function bound to the GPU operator:

    python tools/run_calc_metric.py [--reference] --pred <folder> --data <dataset> [--output metric.json]

`calc_metric.main` keeps doing what it does -- frame_corr.json walk, per-video and overall averages, metric.json
(calc_metric.py:130-233) -- while `calc_metric.calc_metric` (one frame: PNG reads + five metrics, :48-128) is
tcvom_b200.metrics.calc_metric.  The frames run serially in this process (`--n_threads 0`, calc_metric.py:180-183): the
reference's worker pool forks, which a CUDA context does not survive, and is not needed at 23 us per frame.
--reference runs the untouched function the same way (the comparison arm)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    argv = sys.argv[1:]
    reference = "--reference" in argv
    argv = [a for a in argv if a != "--reference"]
    from baseline import ref_env
    ref_env.activate(cpu=reference)
    import calc_metric as cm                     # the reference module
    if not reference:
        import tcvom_b200
        cm.calc_metric = tcvom_b200.metrics.calc_metric
    sys.argv = [os.path.join(ref_env.ref_dir(), "calc_metric.py")] + argv + ["--n_threads", "0"]
    cm.main(cm.parser())


if __name__ == "__main__":
    main()
