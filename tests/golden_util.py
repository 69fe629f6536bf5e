"""Helpers around tests/golden/*.tar.gz (fixtures made by tests/make_golden.py from the reference binaries)."""
import os
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
NAMES = ("g1", "g2", "g3", "g4", "g5", "g6", "g7", "g8", "g9")


def extract(name, dest):
    with tarfile.open(os.path.join(GOLDEN, name + ".tar.gz")) as t:
        t.extractall(dest)
    return os.path.join(dest, name)


def expected(case, mode):
    d = os.path.join(case, "expected", mode)
    return {f: open(os.path.join(d, f), "rb").read() for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt")}


def parse_counts(path):
    """{(gapstart, Lg, round): ndarray[Lg,5]} -- first occurrence of each key."""
    out, key, rows = {}, None, []

    def flush():
        if key is not None and key not in out:
            out[key] = np.array(rows, dtype=np.float64).reshape(-1, 5)

    with open(path) as f:
        for line in f:
            if line.startswith("CALL"):
                flush()
                t = line.split()
                key, rows = (int(t[2]), int(t[4]), int(t[6])), []
            else:
                rows.append([float(x) for x in line.split()])
    flush()
    return out


def model_lines(path):
    return [l for l in open(path).read().split("\n") if l and not l.startswith("gapprob")]
