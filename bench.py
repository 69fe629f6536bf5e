#!/usr/bin/env python
"""bench.py -- Figbird gap-fill hot path on B200: read x offset placements scored per second, gap-fill wall time.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (torchrun for N>1, one rank per GPU) prints ONE
JSON line on rank 0.  A *step* = one pass of the hot path over one batch = the two FillGaps invocations of a
RunFigbird.sh iteration pair (partial mode, then unmapped mode; RunFigbird.sh:352) on ONE synthetic draft, through the
drop-in C-ABI call fb_fillgaps_main with inputs on the host (files), host<->device copies inside the timed region.

Workload (default `c4`, the configuration BASELINE.json quotes its metric on: "100 Mbp draft, 10k gaps, 2x150 bp at 500 bp
insert, 40x coverage, 1/2/4/8 B200"): the gaps of that ONE draft are sharded over the N GPUs (strong scaling) by the
product's own cost-balanced sharding (fb_fillgaps.cpp), driven in-process from rank 0 with FIGBIRD_GPUS=0..N-1; ranks
1..N-1 hold their GPU's place in the torchrun job and meet rank 0 at the barriers (host-side, gloo: a NCCL barrier would
spin on the GPUs rank 0 is timing).  No collective on the data path (gaps are independent).
`--workload c2` = BASELINE configs[1] (4.6 Mbp, 500 gaps, 2x100 @ 200, 50x, the whole 50x SAM).

  value      reference-equivalent pass-1 placements (what the reference's scan consumes: validated against the reference's own
             counters, tests/golden/placements.json) / CUDA-event time of the engine kernels (busiest GPU) of the timed steps.
  e2e.value  the same placements / wall time of the timed steps (files in, files out) -- the headline vs --impl reference.
  roofline   FP64-issue bound (neither HBM- nor tensor-bound): algorithmic FP64 operations (4 per pass-1 base
             term, 1 per pass-2 base term, counted as the reference evaluates them) of everything the device scored / kernel
             time, against the no-FMA FP64 rate measured on this GPU by fb_microbench_fp64.
  cpu_baseline / --impl reference: the reference's own FillGaps + worker (oracle/_ref, as-shipped -O0 worker through the g++
             shim, numthreads = host cores) on a bounded sample of the same workload; placements counted by the reference's
             own counters (oracle/count_patch.awk), not by this repository's replay.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (fbgen args, readlen, insert)
    # BASELINE configs[3]: 100 Mbp, 10 k gaps of 10-2000 bp, 2x150 bp @ 500 bp, 40x.  SAM restricted to the pairs within 900 bp
    # of a gap (everything Preprocess can bucket for a gap: maxDistance is 500 / 575) plus 300 k pairs elsewhere for the model
    "c4": ({"genome": 100000000, "scaffolds": 200, "gaps": 10000, "gapmin": 10, "gapmax": 2000, "cov": 40, "sd": 50, "near": 900, "model-pairs": 300000, "threads": 0}, 150, 500),
    # BASELINE configs[1]: 4.6 Mbp, 500 gaps of 10-500 bp, 2x100 bp @ 200 bp, 50x -- the whole 50x SAM (2.3 M myout.sam lines)
    "c2": ({"genome": 4600000, "scaffolds": 10, "gaps": 500, "gapmin": 10, "gapmax": 500, "cov": 50, "sd": 20}, 100, 200),
    # round-1 operating points (SAM restricted to near-gap pairs + 300 k model pairs)
    "c2n": ({"genome": 4600000, "scaffolds": 10, "gaps": 500, "gapmin": 10, "gapmax": 500, "cov": 50, "sd": 20, "near": 700, "model-pairs": 300000}, 100, 200),
    "c4s": ({"genome": 5000000, "scaffolds": 10, "gaps": 500, "gapmin": 10, "gapmax": 2000, "cov": 40, "sd": 50, "near": 1500, "model-pairs": 300000}, 150, 500),
    # BASELINE configs[2]: the C2 draft with two libraries, the schedule of RunFigbird.sh:552-565 -- partial mode on the 200 bp fragment
    # library, unmapped mode on the 3500 bp jump library (maxDistance 4025).  Same seed and gap spacing => the same draft for both.
    "c3": {"partial": ({"genome": 4600000, "scaffolds": 10, "gaps": 500, "gapmin": 10, "gapmax": 500, "cov": 50, "sd": 20, "minsep": 4175, "near": 700, "model-pairs": 300000}, 100, 200),
           "unmapped": ({"genome": 4600000, "scaffolds": 10, "gaps": 500, "gapmin": 10, "gapmax": 500, "cov": 20, "sd": 350, "minsep": 4175}, 100, 3500)},
    # BASELINE configs[4] regime at single-GPU size: gaps of 10-5000 bp of which 5 % are negative-overlap gaps, two libraries (2x150 @ 500
    # at 40x for partial mode, 2x150 @ 3500 at 15x for unmapped mode); 600 gaps on 9 Mbp instead of 50 k gaps on 250 Mbp
    "c5s": {"partial": ({"genome": 9000000, "scaffolds": 6, "gaps": 600, "gapmin": 10, "gapmax": 5000, "cov": 40, "sd": 50, "minsep": 4250, "negfrac": 0.05, "near": 900, "model-pairs": 300000}, 150, 500),
            "unmapped": ({"genome": 9000000, "scaffolds": 6, "gaps": 600, "gapmin": 10, "gapmax": 5000, "cov": 15, "sd": 350, "minsep": 4250, "negfrac": 0.05}, 150, 3500)},
    "c1": ({"genome": 1000000, "scaffolds": 4, "gaps": 50, "gapmin": 10, "gapmax": 500, "cov": 30, "sd": 20, "near": 700, "model-pairs": 150000}, 100, 200),
    "tiny": ({"genome": 80000, "scaffolds": 1, "gaps": 8, "gapmin": 5, "gapmax": 300, "cov": 30, "sd": 20}, 100, 200),
}
SEEDS = {"c4": 104, "c2": 102, "c3": 103, "c5s": 105, "c2n": 102, "c4s": 102, "c1": 101, "tiny": 7}
# bounded samples of a workload for the CPU reference: same generator parameters (gap lengths, reads, coverage, density of gaps)
# on a shorter draft.  (gaps for the reference arm, gaps for the in-line cpu_baseline of our arm)
# (a C4 gap is ~2e7 placements = about a core-minute of the as-shipped worker, the heaviest ones several: samples stay small so that
#  the reference arm -- one as-shipped step and one steady-state step of one gap per worker -- ends in about five minutes on 16 cores;
#  the arm and the in-line cpu_baseline run the same seeded sample when their sizes agree)
SAMPLE_GAPS = {"c4": (16, 16), "c4s": (16, 16), "c2": (500, 32), "c3": (16, 16), "c5s": (16, 16), "c2n": (500, 32), "c1": (50, 16), "tiny": (8, 8)}


def gen_of(workload):
    """Generator options of a workload (of its partial-mode library when it has two)."""
    w = WORKLOADS[workload]
    return w["partial"][0] if isinstance(w, dict) else w[0]


def sample_spec(workload, gaps):
    if isinstance(WORKLOADS[workload], dict):
        return {m: _sample_spec(WORKLOADS[workload][m], gaps, gen_of(workload)["gaps"]) for m in ("partial", "unmapped")}
    return _sample_spec(WORKLOADS[workload], gaps, WORKLOADS[workload][0]["gaps"])


def _sample_spec(spec, gaps, all_gaps):
    gen, readlen, insert = spec
    g = dict(gen)
    per_gap = g["genome"] / g["gaps"]
    nsc = max(1, min(g["scaffolds"], gaps // 16))
    g.update({"gaps": gaps, "scaffolds": nsc, "genome": int(per_gap * gaps)})
    g.pop("threads", None)
    if "model-pairs" in g:
        g["model-pairs"] = max(20000, int(g["model-pairs"] * gaps / all_gaps))
    return (g, readlen, insert)


def rank_info():
    from figbird_b200.ranks import rank_info as ri
    return ri()


def prepare_case(path, spec, seed):
    """Returns the case directory; for a two-library workload a dict {mode: directory of that mode's library}."""
    import fbcase as fc
    if isinstance(spec, dict):
        cases = {m: prepare_case(os.path.join(path, m + "_lib"), spec[m], seed) for m in ("partial", "unmapped")}
        a, b = (open(os.path.join(cases[m], "draft.fa"), "rb").read() for m in ("partial", "unmapped"))
        if a != b:
            raise RuntimeError("two-library workload: the two libraries were generated on different drafts")
        # RunFigbird.sh keeps ONE Gaps/ directory: the unmapped pass finds the partial files the fragment library left there
        pg, ug = os.path.join(cases["partial"], "partial", "Gaps"), os.path.join(cases["unmapped"], "unmapped", "Gaps")
        marker = os.path.join(ug, ".partial_from_fragment_library")
        if not os.path.exists(marker):
            for f in os.listdir(pg):
                dst = os.path.join(ug, f)
                if os.path.lexists(dst):
                    os.remove(dst)
                os.link(os.path.join(pg, f), dst)
            open(marker, "w").close()
        return cases
    gen, readlen, insert = spec
    if os.path.exists(os.path.join(path, "params.txt")):
        return path
    shutil.rmtree(path, ignore_errors=True)
    g = dict(gen); g["seed"] = seed
    fc.make_case(path, g, readlen=readlen, insert=insert)
    for f in ("result1.sam", "result2.sam", "truth.fa"):
        try:
            os.remove(os.path.join(path, f))
        except OSError:
            pass
    return path


class ClockSampler(threading.Thread):
    def __init__(self, indices):
        super().__init__(daemon=True)
        self.indices, self.samples, self.reasons, self.stop_flag, self.maxmhz = list(indices), [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", ",".join(map(str, self.indices)), "--query-gpu=" + q, "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                for line in o.split("\n"):
                    f = [x.strip() for x in line.split(",")]
                    if len(f) >= 7:
                        if float(f[6] or 0) > 0:     # under load
                            self.samples.append(float(f[0]))
                        self.maxmhz = float(f[1])
                        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                            if v.lower().startswith("active"):
                                self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.maxmhz, "reasons": sorted(self.reasons), "samples": len(s)}


def run_step_ours(case, workdir, metrics):
    """partial + unmapped FillGaps through fb_fillgaps_main; returns summed metrics."""
    import fbcase as fc
    from figbird_b200 import capi
    tot = {}
    for mode in ("partial", "unmapped"):
        mcase = case[mode] if isinstance(case, dict) else case
        tmp = os.path.join(workdir, mode, "Temp")
        os.makedirs(tmp, exist_ok=True)
        for f in ("gapInfo.txt", "stat.txt", "stat2.txt"):
            shutil.copy(os.path.join(mcase, mode, "Temp", f), os.path.join(tmp, f))
        mpath = os.path.join(workdir, "metrics_%s.json" % mode)
        os.environ["FIGBIRD_METRICS"] = mpath
        rc = capi.fillgaps(fc.fillgaps_argv(mcase, mode, tmp, threads=os.cpu_count() or 1))
        if rc != 0:
            raise RuntimeError("fb_fillgaps_main returned %d" % rc)
        m = json.load(open(mpath))
        for k, v in m.items():
            if isinstance(v, (int, float)):
                tot[k] = tot.get(k, 0) + v
            elif isinstance(v, list):
                tot[k] = [a + b for a, b in zip(tot.get(k, [0] * len(v)), v)]
        tot["engine"] = m["engine"]
    metrics.append(tot)
    return tot


def count_reference_placements(case, cores, key=None):
    """The reference's own count of pass-1 placements on a case: counter-instrumented worker (oracle/count_patch.awk), all gaps
    split over `cores` worker processes; untimed.  The count of a seeded sample is a constant: tests/golden/bench_sample_counts.json
    keeps the ones computed where the reference was built (same counting run, tools/make_sample_counts.py), so a bench run does
    not spend minutes of box time on recounting."""
    import fbcase as fc
    if key is not None:
        try:
            known = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_sample_counts.json")))
            if key in known:
                return int(known[key])
        except Exception:
            pass
    total = 0
    for mode in ("partial", "unmapped"):
        mcase = case[mode] if isinstance(case, dict) else case
        cdir = os.path.join(mcase, "count_" + mode)
        shutil.rmtree(cdir, ignore_errors=True); os.makedirs(cdir)
        os.environ["FB_COUNT_DIR"] = cdir
        try:
            fc.run_reference_workers_parallel(mcase, mode, cores, "figbird_worker_count")
        finally:
            os.environ.pop("FB_COUNT_DIR", None)
        total += sum(int(open(os.path.join(cdir, f)).read()) for f in os.listdir(cdir))
    return total


def mode_case(case, mode):
    return case[mode] if isinstance(case, dict) else case


def run_step_reference(case, threads):
    import fbcase as fc
    return sum(fc.run_reference(mode_case(case, mode), mode, threads=threads, worker="figbird_worker_O0")["seconds"] for mode in ("partial", "unmapped"))


def steady_reference(case, cores, worker):
    import fbcase as fc
    return sum(fc.run_reference_workers_parallel(mode_case(case, mode), mode, cores, worker) for mode in ("partial", "unmapped"))


def cpu_baseline(workload, base, cores, unit):
    """The reference's FillGaps as shipped on a small seeded sample of the workload, all host cores (the in-line `cpu_baseline`)."""
    ng = SAMPLE_GAPS[workload][1]
    sample = prepare_case(os.path.join(base, "%s_sample%d" % (workload, ng)), sample_spec(workload, ng), SEEDS[workload] + 2000)
    placements = count_reference_placements(sample, cores, "%s_sample%d_seed%d" % (workload, ng, SEEDS[workload] + 2000))
    secs = run_step_reference(sample, cores)
    out = {"value": placements / secs, "unit": unit, "cores": cores, "kind": "reference",
           "sample": "%d-gap sample of the %s workload (same generator parameters), FillGaps partial+unmapped, as-shipped -O0 worker, numthreads=%d, %.1f s; placements by the reference's own counters" % (ng, workload, cores, secs)}
    if os.environ.get("FB_BENCH_FULL_BASELINE"):      # the steady-state and -O2 flavours (minutes more; the reference arm reports the former)
        steady = steady_reference(sample, cores, "figbird_worker_O0")
        tuned = steady_reference(sample, cores, "figbird_worker_O2")
        out["steady_state"] = {"value": placements / steady, "seconds": steady, "note": "same sample, %d as-shipped -O0 worker processes started together (no driver sleep, no run-time compile)" % min(cores, ng)}
        out["tuned"] = {"value": placements / tuned, "seconds": tuned, "note": "same, worker built -O2 -D_FORTIFY_SOURCE=0 (plain -O2 aborts, SURVEY 5)"}
    return out


def gaps_of(case):
    return len(open(os.path.join(mode_case(case, "partial"), "partial", "Temp", "gapInfo.txt")).readlines())


def assemble_line(a, world, metrics, dt_max, clk, mb, sm_count, ngaps, work, unit, metric, config, t_prep):
    """The JSON line of our arm from the per-call metrics of the timed steps (everything that needs no GPU: tests/test_bench_cpu.py
    runs it on recorded metrics)."""
    S = max(a.steps, 1)
    dev_ms = sum(m["device_ms"] for m in metrics)                 # busiest GPU per FillGaps call, summed
    per_gpu = [sum(m["device_ms_per_gpu"][i] for m in metrics) for i in range(len(metrics[0]["device_ms_per_gpu"]))] if metrics else []
    dev_p1 = sum(m["dev_placements_p1"] for m in metrics); dev_p2 = sum(m["dev_placements_p2"] for m in metrics)
    ref_p1 = sum(m["ref_placements_p1"] for m in metrics)
    terms = sum(m["dev_base_terms"] for m in metrics)
    launches = sum(m["kernel_launches"] for m in metrics)
    h2d = sum(m["h2d_bytes"] for m in metrics); d2h = sum(m["d2h_bytes"] for m in metrics)
    # ---- roofline of the dominant (only) kernel fb_em_kernel: FP64 pipe, no-FMA ceiling (mb = fb_microbench_fp64 on this GPU)
    # algorithmic FP64 ops = 4 per pass-1 base term + 1 per pass-2 base term.  base_terms counts both passes; pass 1 and pass 2 score
    # the same (read, offset) pairs, so split by placements.  Kernel time = sum over GPUs of each GPU's busy time (union of its kernel
    # intervals): the figure is per GPU.
    t1 = terms * dev_p1 / max(dev_p1 + dev_p2, 1); t2 = terms - t1
    flops = 4.0 * t1 + 1.0 * t2
    busy_ms = sum(per_gpu) if per_gpu else dev_ms
    ach = flops / (busy_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    lane1 = sum(m.get("lane_steps_p1", 0) for m in metrics); lane2 = sum(m.get("lane_steps_p2", 0) for m in metrics)
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except Exception:
        pass
    dom = ncu.get("unmapped", {})
    smem_peak = 128.0 * sm_count * (clk["sm_mhz"] or 1965.0) * 1e6 / 1e9
    roof = {"bound": "fp64-issue",
            "bound_note": "neither HBM- nor tensor-bound (SURVEY.md 8d): FP64 issue without FMA (4 separately rounded ops per pass-1 term as the reference computes them); "
                          "co-limited by the shared-memory crossbar: one 16-byte table entry per gap-row term (DESIGN.md 3)",
            "achieved": ach, "peak": mb["dmul_tinstr_s"], "unit": "TFLOP/s", "frac": ach / mb["dmul_tinstr_s"] if mb["dmul_tinstr_s"] else None,
            "per": "GPU (algorithmic flop of all GPUs / summed per-GPU kernel time)",
            "traffic": dom.get("dram_bytes_per_launch"),
            "traffic_note": ("dram__bytes_read+write of the dominant fb_em_kernel launch (ncu --set full, profiles/ncu_summary.json: %s, grid %s, %.1f ms); algorithmic HBM bytes of that launch ~= inputs once + result arena"
                             % (dom.get("file"), dom.get("grid"), dom.get("duration_ms", 0.0))) if dom else None,
            "peak_source": "fb_microbench_fp64 on this GPU (DMUL chains, 1 flop/instr); DFMA rate %.1f TFLOP/s" % mb["dfma_tflops"],
            "algorithmic": "4 FP64 ops per pass-1 base term, 1 per pass-2 base term (SURVEY.md 8d); %.3e base terms per step" % (terms / S),
            "kernel_ms_per_launch": busy_ms / max(launches, 1),
            "executed_vs_algorithmic": {"pass1_lane_steps": lane1, "pass2_lane_steps": lane2, "algorithmic_terms_pass1": t1, "algorithmic_terms_pass2": t2,
                                        "note": "flank terms come from the per-gap cache and pass 2 is pruned, so the kernel walks fewer terms than the reference evaluates"},
            "smem": {"achieved_gbs": 16.0 * lane1 / (busy_ms * 1e-3) / 1e9, "peak_gbs": smem_peak, "frac": 16.0 * lane1 / (busy_ms * 1e-3) / 1e9 / smem_peak,
                     "note": "pass-1 walk only: 16 B (LDS.128 of {P, E-P}) per executed gap-row lane step / kernel time, against 128 B/clk/SM x SMs x SM clock"},
            "ncu": {k: dom.get(k) for k in ("issue_active_pct", "smem_wavefronts_pct", "fp64_pipe_pct", "alu_pipe_pct", "lsu_pipe_pct", "warps_active_pct", "stall_barrier")} if dom else None,
            # the busiest unit of the dominant launches according to ncu (committed capture, not this run): the shared-memory pipe
            "smem_pipe": ({"achieved": dom.get("smem_wavefronts_pct"), "peak": 100.0, "unit": "%% of peak shared-memory wavefronts (ncu l1tex__data_pipe_lsu_wavefronts_mem_shared, dominant launch of profiles/%s)" % dom.get("file"),
                           "frac": (dom.get("smem_wavefronts_pct") or 0.0) / 100.0} if dom else None),
            "hbm": {"achieved_gbs": (h2d + d2h) / (busy_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "note": "algorithmic HBM bytes ~= result arena + inputs; tables live in shared memory"}}
    host = {k: sum(m[k] for m in metrics) / S for k in ("t_load", "t_model", "t_model_wait", "t_prepare", "t_fill", "t_write", "t_ctx_upload", "t_workers", "t_engine_calls", "cpu_workers") if metrics and k in metrics[0]}
    wall = dt_max / S
    parts = {"kernels (busiest GPU)": dev_ms * 1e-3 / S,
             "GPU idle inside the fill phase (host replay between engine calls, tail, imbalance across GPUs)": max(0.0, host.get("t_fill", 0) - dev_ms * 1e-3 / S),
             "serial host phases (load, prepare, write, python glue)": max(0.0, wall - host.get("t_fill", 0))}
    limiter = max(parts, key=parts.get)
    line = {"metric": metric, "value": ref_p1 / (dev_ms * 1e-3), "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * wall, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clk,
            "e2e": {"value": ref_p1 / dt_max, "unit": unit, "h2d_bytes_per_step": h2d / S, "d2h_bytes_per_step": d2h / S,
                    "note": "through fb_fillgaps_main: files -> model -> per-gap control on host fibers -> engine -> files; reference-equivalent pass-1 placements / wall",
                    "gaps_per_s": ngaps * S / dt_max, "gaps": ngaps, "seconds_per_step": wall,
                    "seconds_per_step_breakdown": parts, "limiter": limiter,
                    "host_seconds_per_step": host, "device_ms_per_gpu_per_step": [x / S for x in per_gpu]},
            "gpu_launches": int(launches), "placements": {"reference_equivalent_p1": ref_p1, "device_p1": dev_p1, "device_p2": dev_p2,
                                                          "note": "device counts include speculative candidates past the reference's early exits"},
            "value_device_placements": dev_p1 / (dev_ms * 1e-3), "data_prep_s": t_prep, "roofline": roof}
    # at-size parity: the gapout.txt lines of the last step against the reference worker's lines for a seeded sample of the gaps
    # (tests/golden/<workload>_sample.json, made by tools/make_sample_expect.py where the reference binaries run)
    exp_path = os.path.join(ROOT, "tests", "golden", "%s_sample.json" % a.workload)
    if os.path.exists(exp_path):
        exp = json.load(open(exp_path))
        ps = {"gaps_sampled": len(exp["gaps"]), "of": ngaps, "reference": "oracle/_ref %s on the same seeded case" % exp.get("worker", "worker")}
        for mode in ("partial", "unmapped"):
            ours = {}
            for ln in open(os.path.join(work, mode, "Temp", "gapout.txt")):
                ours[ln.split("\t", 1)[0]] = ln
            ps["identical_" + mode] = sum(1 for g, ln in exp[mode].items() if ours.get(g) == ln)
            bad = [int(g) for g, ln in exp[mode].items() if ours.get(g) != ln]
            if bad:
                ps["different_" + mode] = sorted(bad)[:20]
        line["parity_sample"] = ps
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FB_BENCH_WORKLOAD", "c4"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print at the C level (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank, world, local = rank_info()
    os.environ["FIGBIRD_QUIET"] = "1"
    import fbcase as fc
    cores = os.cpu_count() or 1
    base = os.environ.get("FB_BENCH_DIR", "/tmp/fb_bench")
    unit = "placements/s"
    metric = "read x offset placements scored per second (pass 1, as the reference counts them); gap-fill of synthetic %s" % a.workload
    def describe(spec):
        return "%s, readlen %d, insert %d" % (json.dumps({k: v for k, v in spec[0].items() if k != "threads"}, sort_keys=True), spec[1], spec[2])
    w = WORKLOADS[a.workload]
    wtext = ("partial mode on library {%s}, unmapped mode on library {%s}" % (describe(w["partial"]), describe(w["unmapped"]))) if isinstance(w, dict) else describe(w)
    config = {"workload": "%s: %s; step = FillGaps partial + unmapped (RunFigbird.sh:352) on one draft" % (a.workload, wtext),
              "l2": "flushed between steps (256 MiB device memset per GPU); the per-step inputs are far larger than L2",
              "sharding": "ONE draft; its gaps are sharded cost-balanced over the N GPUs in-process from rank 0 (FIGBIRD_GPUS=0..N-1); strong scaling, no collective on the path"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        if not fc.have_reference():
            emit({"impl": "reference", "unavailable": "oracle/_ref not built (no /root/reference at build time)"})
            return 0
        ng = SAMPLE_GAPS[a.workload][0]
        whole = ng >= gen_of(a.workload)["gaps"]
        if whole:
            sample = prepare_case(os.path.join(base, a.workload), WORKLOADS[a.workload], SEEDS[a.workload])
        else:
            sample = prepare_case(os.path.join(base, "%s_sample%d" % (a.workload, ng)), sample_spec(a.workload, ng), SEEDS[a.workload] + 2000)
        # the reference's own pass-1 counters (oracle/count_patch.awk): a constant of the seeded sample, read from the fixture when it
        # is there, else counted now (-O2 flavour, untimed -- which then also serves as the warm-up run)
        t_count0 = time.perf_counter()
        placements = count_reference_placements(sample, cores, None if whole else "%s_sample%d_seed%d" % (a.workload, ng, SEEDS[a.workload] + 2000))
        counted_now = time.perf_counter() - t_count0 > 1.0
        # every step is the same deterministic CPU job; a wall-clock budget bounds the arm whatever K the caller asks for
        budget = float(os.environ.get("FB_REF_BUDGET_S", "420"))
        t, done = 0.0, 0
        while done < max(a.steps, 1) and (done == 0 or t + t / done <= budget):
            t += run_step_reference(sample, cores); done += 1
        steady = steady_reference(sample, cores, "figbird_worker_O0")
        v = placements * done / t
        what = ("the whole %s workload" % a.workload) if whole else (
            "%d-gap sample of the %s workload (same generator parameters, %d bp draft)" % (ng, a.workload, int(gen_of(a.workload)["genome"] / gen_of(a.workload)["gaps"] * ng)))
        cfg = dict(config); cfg["workload"] = config["workload"] + " -- reference arm ran: " + what
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": a.gpus, "steps": done, "warmup": 1 if counted_now else 0, "steps_requested": a.steps, "warmup_requested": a.warmup,
                "ms_per_step": 1e3 * t / done, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "reference",
                                 "sample": "%s, FillGaps partial+unmapped as shipped (run-time-compiled -O0 worker via the g++ shim, numthreads=%d, 1 s sleep per worker start: FillGaps.cpp:675), %.1f s per step, %d gaps, %d pass-1 placements counted by the reference's own counters (oracle/count_patch.awk)" % (what, cores, t / done, gaps_of(sample), placements),
                                 "steady_state": {"value": placements / steady, "seconds": steady, "note": "same sample, %d as-shipped -O0 worker processes started together (no driver sleep, no run-time compile, round-robin gap split)" % min(cores, gaps_of(sample))}},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "gaps_per_s": gaps_of(sample) * done / t}, "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    torch.cuda.set_device(local)
    dist = None; host_group = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo", timeout=datetime.timedelta(hours=2))      # host-side barriers: nothing spins on a GPU

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(group=host_group)
        torch.cuda.synchronize()

    ngpu = world
    os.environ["FIGBIRD_GPUS"] = ",".join(str(i) for i in range(ngpu))
    case = os.path.join(base, a.workload)
    work = os.path.join(base, "work_%s" % a.workload)
    t_prep0 = time.perf_counter()
    if rank == 0:
        case = prepare_case(case, WORKLOADS[a.workload], SEEDS[a.workload])
    t_prep = time.perf_counter() - t_prep0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    metrics, warm = [], []
    sampler = None
    barrier()
    if rank == 0:
        for _ in range(max(a.warmup, 0)):
            flush.zero_()
            run_step_ours(case, work, warm)
        sampler = ClockSampler(range(ngpu)); sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        flush.zero_()          # every rank flushes its own GPU's L2
        torch.cuda.synchronize()
        if rank == 0:
            run_step_ours(case, work, metrics)
        if dist is not None:
            dist.barrier(group=host_group)
    barrier()
    dt = time.perf_counter() - t0
    if sampler is not None:
        sampler.stop_flag = True; sampler.join(timeout=2)
    from figbird_b200.ranks import reduce_counters
    mx, _ = reduce_counters({"dt": dt}, dist, "cuda")
    dt_max = mx["dt"]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    from figbird_b200 import capi
    eng = capi.Engine(0)
    mb = eng.microbench_fp64(); eng.close()
    line = assemble_line(a, world, metrics, dt_max, sampler.summary(), mb, torch.cuda.get_device_properties(0).multi_processor_count, gaps_of(case), work, unit, metric, config, t_prep)
    if world == 1 and not a.no_cpu_baseline and fc.have_reference():
        try:
            line["cpu_baseline"] = cpu_baseline(a.workload, base, cores, unit)
        except Exception as e:      # the baseline is a side measurement: its failure must not cost the bench line
            line["cpu_baseline"] = {"value": None, "unit": unit, "cores": cores, "kind": "reference", "sample": "failed: %s" % str(e)[:300]}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
