"""``python -m gdtk_b200 --run --job=<name> [--dir=<job directory>]``: the run stage of a prepared job on the
GPU, with the command-line spelling of ``e4shared --run --job=<name>`` (src/eilmer/main.d).  Preparation
(``--prep``) and post-processing (``--post``) stay with the reference's tools; this reads what the former wrote
and writes what the latter reads (gdtk_b200/job.py)."""
import argparse
import sys
import time


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m gdtk_b200")
    ap.add_argument("--run", action="store_true", help="run the simulation over time")
    ap.add_argument("--job", required=True, help="file names are built from this string")
    ap.add_argument("--dir", default=".", help="job directory (holds config/, grid/, flow/)")
    ap.add_argument("--tindx-start", type=int, default=0, help="solution to start from")
    ap.add_argument("--n-solutions", type=int, default=1, help="how many solutions to write between start and max_time")
    ap.add_argument("--max-wall-clock", type=float, default=None, help="(accepted for compatibility; not enforced)")
    ap.add_argument("--verbosity", type=int, default=1)
    ap.add_argument("--strict-fp", action="store_true", help="FMA-free kernels (bit-comparable with the CPU restatement)")
    args = ap.parse_args(argv)
    if not args.run:
        ap.error("only --run is implemented here; use e4shared for --prep and --post")
    from . import job
    t0 = time.time()
    sim = job.run_job(args.dir, args.job, tindx_start=args.tindx_start, n_solutions=args.n_solutions, strict_fp=args.strict_fp)
    if args.verbosity > 0:
        # the line the reference's test scripts look for (e.g. cone20-test.rb:27-31)
        print(f"Step= {sim.step} final-t= {sim.time:.6e} dt= {sim.dt_global:.3e} WC= {time.time() - t0:.1f}")
        print(f"Done simulation, {sim.kernel_launches()} kernel launches.")
    sim.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
