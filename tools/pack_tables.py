#!/usr/bin/env python3
"""Pack the reference's exponential-sum tables into one binary file.

Input (read-only, build container only):
  /root/reference/coefficients_data/output_data/tabelle_complete.csv
      122 rows (R = condition-number bound) x 63 columns (rank t): the max-norm
      error of  sum_j omega_j exp(-alpha_j x)  against 1/x on [1, R]
      (read by approximation.jl:44-54; looked up by approximation.jl:65-84)
  /root/reference/coefficients_data/1_xk%02d.%d_%d  (t, first digit, order)
      2t non-blank lines "value {name}": omega[1..t] then alpha[1..t]
      (read by approximation.jl:119-147)

Output: tensorkrylov.jl_b200/data/expsum_tables.bin, little-endian:

  char   magic[8] = "TKXSUM01"
  int32  nrows, nranks
  double R[nrows]
  double err[nrows][nranks]           (Inf where the CSV says Inf)
  int32  ranks[nranks]                (CSV column headers as integers)
  int32  nfiles, pad
  nfiles x { int32 t, digit, order, pad;  double omega[t];  double alpha[t] }

The values are the CSV/file text parsed to Float64 (correctly rounded), i.e.
what the reference's CSV.read calls produce.  The CSV is packed AS SHIPPED,
including its column "11" whose entries lost their exponents (SURVEY.md
section 8c): rank parity with the reference depends on that quirk.
"""
import csv
import os
import re
import struct
import sys

REF = os.environ.get("TK_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tensorkrylov.jl_b200", "data", "expsum_tables.bin")


def main():
    cdir = os.path.join(REF, "coefficients_data")
    with open(os.path.join(cdir, "output_data", "tabelle_complete.csv")) as f:
        rows = list(csv.reader(f))
    header = rows[0]
    assert header[0] == "R"
    ranks = [int(c) for c in header[1:]]
    R = []
    err = []
    for row in rows[1:]:
        if not row:
            continue
        R.append(float(row[0]))
        err.append([float(x) for x in row[1:]])
        assert len(err[-1]) == len(ranks)

    files = []
    pat = re.compile(r"^1_xk(\d\d)\.(\d+)_(\d+)$")
    for name in sorted(os.listdir(cdir)):
        m = pat.match(name)
        if not m:
            continue
        t, digit, order = int(m.group(1)), int(m.group(2)), int(m.group(3))
        vals = []
        with open(os.path.join(cdir, name)) as f:
            for line in f:
                if line.strip():
                    vals.append(float(line.split("{")[0]))
        assert len(vals) == 2 * t, (name, len(vals))
        files.append((t, digit, order, vals[:t], vals[t:]))

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "wb") as f:
        f.write(b"TKXSUM01")
        f.write(struct.pack("<ii", len(R), len(ranks)))
        f.write(struct.pack("<%dd" % len(R), *R))
        for e in err:
            f.write(struct.pack("<%dd" % len(e), *e))
        f.write(struct.pack("<%di" % len(ranks), *ranks))
        f.write(struct.pack("<ii", len(files), 0))
        for t, digit, order, om, al in files:
            f.write(struct.pack("<iiii", t, digit, order, 0))
            f.write(struct.pack("<%dd" % t, *om))
            f.write(struct.pack("<%dd" % t, *al))
    print("rows", len(R), "ranks", len(ranks), "coefficient files", len(files),
          "bytes", os.path.getsize(OUT))


if __name__ == "__main__":
    sys.exit(main())
