"""GPU-box check: product `fillgaps` (CUDA) vs the reference binaries (oracle/_ref) on a fresh synthetic case."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fbcase as fc  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 31
    ngaps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    genome = int(sys.argv[3]) if len(sys.argv) > 3 else 80000
    case = "/tmp/fb_e2e_%d" % seed
    fc.make_case(case, {"genome": genome, "gaps": ngaps, "gapmin": 5, "gapmax": 500, "seed": seed, "cov": 30})
    ok = True
    for mode in ("partial", "unmapped"):
        r = fc.run_reference(case, mode, threads=max(1, min(ngaps, os.cpu_count() or 1)))
        o = fc.run_ours(case, mode, fc.product_exe(), extra_env={"FIGBIRD_METRICS": os.path.join(case, "metrics_%s.json" % mode)})
        for f in ("gapout.txt", "filledContigs.fa", "Ncount.txt"):
            same = r[f] == o[f]
            ok &= same
            print(mode, f, "SAME" if same else "DIFF")
        dr, do = fc.draw_by_gap(r["draw.txt"]), fc.draw_by_gap(o["draw.txt"])
        print(mode, "draw.txt per gap", "SAME" if dr == do else "DIFF")
        print(mode, "ref %.2fs ours %.2fs" % (r["seconds"], o["seconds"]), open(os.path.join(case, "metrics_%s.json" % mode)).read().strip())
        if r["gapout.txt"] != o["gapout.txt"]:
            for a, b in zip(r["gapout.txt"].decode().split("\n"), o["gapout.txt"].decode().split("\n")):
                if a != b:
                    print("  REF ", a[:160]); print("  OURS", b[:160])
    print("E2E", "OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
