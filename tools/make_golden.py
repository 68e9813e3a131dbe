#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's stored experiment results.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box).  The outputs are small and committed.

Source of the data: experiments/data/reproduction_data/{laplace_new,nonsym_new}
and experiments/data/eigenvalues_data/dzero.  They are Julia `Serialization`
dumps of the `Experiment` struct (experiments/experiment_common.jl:41-66,
written by serialize_to_file, experiment_common.jl:142-155):

    dims::Vector{Int}, matrixsize, nmax, instance, matrixclass, orth_method,
    rhs_vec::Vector{Vector{Vector{Float64}}},   # per d: d vectors of length n
    conv_vector::Vector{ConvergenceData}        # per d: niterations, iterations,
                                                #   relative_residual_norm,
                                                #   projected_residual_norm,
                                                #   orthogonality_data

Only the 1-d arrays are needed.  In the serialized stream a 1-d array is
    0x15 0x00 <eltype tag> <length> <raw little-endian data>
with eltype tag 0x0e = Float64, 0x08 = Int64, and <length> either
0x31 + int32 or a single byte 0xdf+n for n <= 32.  Arrays appear in struct
field order, which is all this decoder relies on.
"""
import os
import sys

import numpy as np

REF = os.environ.get("TK_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def extract_arrays(path):
    """The product's own decoder of the stream (tensorkrylov.jl_b200/experiments.py::julia_arrays)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import __graft_entry__ as entry
    return entry.load_package().experiments.julia_arrays(open(path, "rb").read())


def decode_experiment(path, n):
    arrs = extract_arrays(path)
    assert arrs[0][1] == "i", "first array must be dims"
    dims = arrs[0][2]
    floats = [(o, a) for o, k, a in arrs if k == "f"]
    ints = [(o, a) for o, k, a in arrs if k == "i"][1:]
    rhs_all = [a for o, a in floats if len(a) == n][: int(dims.sum())]
    offs = np.concatenate([[0], np.cumsum(dims)])
    result = {"dims": dims}
    for i, d in enumerate(dims):
        rhs = rhs_all[offs[i]:offs[i + 1]]
        # the reference replicates one vector over all modes (system.jl:5-11)
        assert all(np.array_equal(rhs[0], r) for r in rhs)
        result[f"rhs_d{d}"] = rhs[0]
        o_iter, iters = ints[i]
        conv = [a for o, a in floats if o > o_iter][:3]
        assert all(len(c) == len(iters) for c in conv)
        result[f"iterations_d{d}"] = iters
        result[f"relres_d{d}"] = conv[0]
        result[f"projres_d{d}"] = conv[1]
        result[f"orth_d{d}"] = conv[2]
    return result


def main():
    os.makedirs(OUT, exist_ok=True)
    jobs = [
        ("laplace_new", "experiments/data/reproduction_data/laplace_new", 200),
        ("nonsym_new", "experiments/data/reproduction_data/nonsym_new", 200),
        ("eigval_dzero", "experiments/data/eigenvalues_data/dzero", 200),
    ]
    for name, rel, n in jobs:
        res = decode_experiment(os.path.join(REF, rel), n)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
        dims = res["dims"]
        print(name, "dims", dims, "iters", [len(res[f"iterations_d{d}"]) for d in dims],
              "final relres", [float(res[f"relres_d{d}"][-1]) for d in dims])


if __name__ == "__main__":
    sys.exit(main())
