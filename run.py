"""Command-line entry, same positional arguments and -d/-k flags as the reference's run.py
(reference: run.py:5-84).  FEM back-end only: elasticity designs (the hot path) and fluid designs.

    python run.py designs/cantilever.json 40 [-d output] [-k 1] [--dtype float64]
"""
import argparse


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__)
    parser.add_argument("design_file", type=argparse.FileType("r"),
                        help="path to a json file where your problem is defined")
    parser.add_argument("N", metavar="element_count", type=int,
                        help="the number of finite elements in a unit length")
    parser.add_argument("-d", "--data_path", default="output",
                        help="the folder where the data output is stored (default: 'output')")
    parser.add_argument("-k", "--skip_multiple", type=int, default=1,
                        help="only store data when the iteration is a multiple of this; "
                        "first and last iteration are always stored (default: 1)")
    parser.add_argument("-n", "--use_neural_network_solver", action="store_true",
                        help="(reference flag) the DEM back-end is not part of this package")
    parser.add_argument("--dtype", choices=("float64", "float32"), default="float64")
    parser.add_argument("--preconditioner", choices=("multigrid", "jacobi"), default="multigrid")
    args = parser.parse_args(argv)
    design_filename = args.design_file.name
    args.design_file.close()
    if args.use_neural_network_solver:
        parser.error("the DEM solver is outside the accelerated path; use the reference for it")

    from FEM_src.solver import FEMSolver

    solver = FEMSolver(args.N, design_filename, args.data_path, args.skip_multiple, dtype=args.dtype,
                       problem_options={"preconditioner": args.preconditioner})
    solver.solve()


if __name__ == "__main__":
    main()
