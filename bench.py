#!/usr/bin/env python
"""bench.py -- Figbird gap-fill hot path on B200: read x offset placements scored per second.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (torchrun for N>1, one rank per GPU) prints ONE
JSON line on rank 0.  A *step* = one pass of the hot path over one batch = the two FillGaps invocations of a
RunFigbird.sh iteration pair (partial mode, then unmapped mode; RunFigbird.sh:352) on the synthetic C2 draft
(BASELINE.json configs[1]: 4.6 Mbp, 500 gaps, 2x100 bp @ 200 bp, 50x), through the drop-in C-ABI call
fb_fillgaps_main with inputs on the host (files), host<->device copies inside the timed region.

  value      device placements (pass 1) / CUDA-event time of the engine kernels of the timed steps
             (batch resident in HBM when each kernel starts); explains the roofline.
  e2e.value  reference-equivalent placements / wall time of the timed steps (the headline vs --impl reference).
  roofline   FP64-pipe bound (the path is neither HBM- nor tensor-bound): algorithmic FP64
             operations (4 per pass-1 base term, 1 per pass-2 base term) / kernel time (union of the kernel
             intervals on the device), against the no-FMA FP64 rate measured on this GPU by fb_microbench_fp64.
             Beside it: the shared-memory view of the pass-1 walk (16 B per executed gap-row term), executed vs
             algorithmic terms, HBM GB/s, and the counters of the dominant launch from the committed ncu capture.
  cpu_baseline / --impl reference: the reference's own FillGaps + worker (oracle/_ref, as-shipped -O0 worker
             through the g++ shim) on a bounded sample of the same workload, all host cores.
N>1: weak scaling -- every rank fills its own C2-sized draft (seed + rank), no collective on the path.
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
    "c2": ({"genome": 4600000, "scaffolds": 10, "gaps": 500, "gapmin": 10, "gapmax": 500, "cov": 50, "sd": 20, "near": 700, "model-pairs": 300000}, 100, 200),
    # BASELINE configs[3] read / gap parameters (2x150 bp @ 500 bp, gaps 10-2000 bp, 40x) at single-GPU size: not the headline, a
    # second operating point (150-base reads, candidates whose tables live in global memory, large-gap rounds)
    "c4s": ({"genome": 5000000, "scaffolds": 10, "gaps": 500, "gapmin": 10, "gapmax": 2000, "cov": 40, "sd": 50, "near": 1500, "model-pairs": 300000}, 150, 500),
    "c1": ({"genome": 1000000, "scaffolds": 4, "gaps": 50, "gapmin": 10, "gapmax": 500, "cov": 30, "sd": 20, "near": 700, "model-pairs": 150000}, 100, 200),
    "tiny": ({"genome": 80000, "scaffolds": 1, "gaps": 8, "gapmin": 5, "gapmax": 300, "cov": 30, "sd": 20}, 100, 200),
}
SAMPLE = ({"genome": 294400, "scaffolds": 2, "gaps": 32, "gapmin": 10, "gapmax": 500, "cov": 50, "sd": 20}, 100, 200)   # same density as c2


def rank_info():
    from figbird_b200.ranks import rank_info as ri
    return ri()


def prepare_case(path, spec, seed):
    import fbcase as fc
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
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.maxmhz = index, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                f = [x.strip() for x in o.split(",")]
                if len(f) >= 6:
                    self.samples.append(float(f[0])); self.maxmhz = float(f[1])
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
        tmp = os.path.join(workdir, mode, "Temp")
        os.makedirs(tmp, exist_ok=True)
        for f in ("gapInfo.txt", "stat.txt", "stat2.txt"):
            shutil.copy(os.path.join(case, mode, "Temp", f), os.path.join(tmp, f))
        mpath = os.path.join(workdir, "metrics_%s.json" % mode)
        os.environ["FIGBIRD_METRICS"] = mpath
        rc = capi.fillgaps(fc.fillgaps_argv(case, mode, tmp, threads=os.cpu_count() or 1))
        if rc != 0:
            raise RuntimeError("fb_fillgaps_main returned %d" % rc)
        m = json.load(open(mpath))
        for k, v in m.items():
            if isinstance(v, (int, float)):
                tot[k] = tot.get(k, 0) + v
        tot["engine"] = m["engine"]
    metrics.append(tot)
    return tot


def run_step_reference(case, threads):
    import fbcase as fc
    secs = 0.0
    for mode in ("partial", "unmapped"):
        r = fc.run_reference(case, mode, threads=threads, worker="figbird_worker_O0")
        secs += r["seconds"]
    return secs


def ref_equivalent_placements(case, workdir):
    """Pass-1 placements the reference scan consumes on this case (counted by our host replay of the same scan)."""
    ms = []
    run_step_ours(case, workdir, ms)
    return ms[0]["ref_placements_p1"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
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
    metric = "read x offset placements scored per second (pass 1), gap-fill of synthetic %s" % a.workload
    config = {"workload": "%s: %s, step = FillGaps partial + unmapped (RunFigbird.sh:352)" % (a.workload, json.dumps(WORKLOADS[a.workload][0], sort_keys=True)),
              "l2": "flushed between steps (256 MiB device memset)", "scaling_note": "weak: each rank fills its own draft (seed+rank)"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        if not fc.have_reference():
            emit({"impl": "reference", "unavailable": "oracle/_ref not built (no /root/reference at build time)"})
            return 0
        sample = prepare_case(os.path.join(base, "sample"), SAMPLE, 1102)
        # placements of the sample, counted once by our replay (needs the GPU library; outside the timed region)
        placements = ref_equivalent_placements(sample, os.path.join(base, "sample_work"))
        # every step is the same deterministic CPU job (~45 s as shipped: the driver sleeps 1 s per worker start), so one warm-up
        # and a wall-clock budget bound the arm to a few minutes whatever K and W the caller asks for
        budget = float(os.environ.get("FB_REF_BUDGET_S", "200"))
        for _ in range(a.warmup if a.warmup < 1 else 1):
            run_step_reference(sample, cores)
        t, done = 0.0, 0
        while done < max(a.steps, 1) and (done == 0 or t + t / done <= budget):
            t += run_step_reference(sample, cores); done += 1
        v = placements * done / t
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "steps_run": done, "ms_per_step": 1e3 * t / done,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "reference",
                                 "sample": "32-gap / 294 kbp sample of the c2 workload (same gap, read and coverage parameters), FillGaps partial+unmapped, as-shipped -O0 worker, numthreads=%d" % cores},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["FIGBIRD_GPUS"] = str(local)
    if world > 1 and "FIGBIRD_HOST_THREADS" not in os.environ:      # ranks of one box share its host cores
        os.environ["FIGBIRD_HOST_THREADS"] = str(max(2, cores // world))
    case = prepare_case(os.path.join(base, "%s_rank%d" % (a.workload, rank)), WORKLOADS[a.workload], 102 + rank)
    work = os.path.join(base, "work_rank%d" % rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    warm = []
    for _ in range(max(a.warmup, 0)):
        flush.zero_()
        run_step_ours(case, work, warm)
    sampler = ClockSampler(local); sampler.start()
    metrics = []
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        flush.zero_()
        run_step_ours(case, work, metrics)
    barrier()
    dt = time.perf_counter() - t0
    sampler.stop_flag = True; sampler.join(timeout=2)

    dev_ms = sum(m["device_ms"] for m in metrics)
    dev_p1 = sum(m["dev_placements_p1"] for m in metrics); dev_p2 = sum(m["dev_placements_p2"] for m in metrics)
    ref_p1 = sum(m["ref_placements_p1"] for m in metrics)
    terms = sum(m["dev_base_terms"] for m in metrics)
    launches = sum(m["kernel_launches"] for m in metrics)
    h2d = sum(m["h2d_bytes"] for m in metrics); d2h = sum(m["d2h_bytes"] for m in metrics)
    from figbird_b200.ranks import reduce_counters
    mx, sm = reduce_counters({"dt": dt, "dev_ms": dev_ms, "dev_p1": dev_p1, "ref_p1": ref_p1, "launches": launches, "terms": terms, "dev_p2": dev_p2}, dist, "cuda")
    dt_max, devms_max = mx["dt"], mx["dev_ms"]
    dev_p1_all, ref_p1_all, launches_all, terms_all, dev_p2_all = sm["dev_p1"], sm["ref_p1"], sm["launches"], sm["terms"], sm["dev_p2"]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel fb_em_kernel: FP64 pipe, no-FMA ceiling
    from figbird_b200 import capi
    eng = capi.Engine(local)
    mb = eng.microbench_fp64(); eng.close()
    # per-rank figures (this rank's kernels): algorithmic FP64 ops = 4 per pass-1 base term + 1 per pass-2 base term.
    # base_terms counts both passes; pass-1 and pass-2 score the same (read, offset) pairs, so split by placements.
    t1 = terms * dev_p1 / max(dev_p1 + dev_p2, 1); t2 = terms - t1
    flops = 4.0 * t1 + 1.0 * t2
    ach = flops / (dev_ms * 1e-3) / 1e12
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
    dom = ncu.get("unmapped", {})      # the unmapped-mode launches carry ~3/4 of the kernel time (profiles/README.md)
    smem_peak = 128.0 * torch.cuda.get_device_properties(local).multi_processor_count * (sampler.summary()["sm_mhz"] or 1965.0) * 1e6 / 1e9
    roof = {"bound": "fp64-issue",
            "bound_note": "neither HBM- nor tensor-bound (SURVEY.md 8d): FP64 issue without FMA (4 separately rounded ops per pass-1 term as the reference computes them); "
                          "co-limited by the shared-memory crossbar: one 16-byte table entry per gap-row term (DESIGN.md 3)",
            "achieved": ach, "peak": mb["dmul_tinstr_s"], "unit": "TFLOP/s",
            "frac": ach / mb["dmul_tinstr_s"] if mb["dmul_tinstr_s"] else None,
            "traffic": dom.get("dram_bytes_per_launch"),
            "traffic_note": "dram__bytes_read+write of the dominant fb_em_kernel launch (ncu --set full, profiles/ncu_summary.json: %s, grid %s, %.1f ms); algorithmic HBM bytes of that launch ~= inputs once + result arena"
                            % (dom.get("file"), dom.get("grid"), dom.get("duration_ms", 0.0)) if dom else None,
            "peak_source": "fb_microbench_fp64 on this GPU (DMUL chains, 1 flop/instr); DFMA rate %.1f TFLOP/s" % mb["dfma_tflops"],
            "algorithmic": "4 FP64 ops per pass-1 base term, 1 per pass-2 base term (SURVEY.md 8d); %.3e base terms per step" % (terms / max(a.steps, 1)),
            "kernel_ms_per_launch": dev_ms / max(launches, 1),
            "executed_vs_algorithmic": {"pass1_lane_steps": lane1, "pass2_lane_steps": lane2, "algorithmic_terms_pass1": t1, "algorithmic_terms_pass2": t2,
                                        "note": "flank terms come from the per-gap cache and pass 2 is pruned, so the kernel walks fewer terms than the reference evaluates"},
            "smem": {"achieved_gbs": 16.0 * lane1 / (dev_ms * 1e-3) / 1e9, "peak_gbs": smem_peak, "frac": 16.0 * lane1 / (dev_ms * 1e-3) / 1e9 / smem_peak,
                     "note": "pass-1 walk only: 16 B (LDS.128 of {P, E-P}) per executed gap-row lane step / kernel time, against 128 B/clk/SM x SMs x SM clock"},
            "issue": ({"achieved": dom.get("issue_active_pct"), "peak": 100.0, "unit": "% of issue slots active (ncu sm__issue_active, dominant launch of the committed capture)",
                       "frac": (dom.get("issue_active_pct") or 0.0) / 100.0} if dom else None),
            "ncu": {k: dom.get(k) for k in ("issue_active_pct", "smem_wavefronts_pct", "fp64_pipe_pct", "alu_pipe_pct", "lsu_pipe_pct", "warps_active_pct", "stall_barrier")} if dom else None,
            "hbm": {"achieved_gbs": (h2d + d2h) / (dev_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "note": "algorithmic HBM bytes ~= result arena + inputs; tables live in shared memory"}}
    line = {"metric": metric, "value": dev_p1_all / (devms_max * 1e-3), "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * dt_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": sampler.summary(),
            "e2e": {"value": ref_p1_all / dt_max, "unit": unit, "h2d_bytes_per_step": h2d / a.steps, "d2h_bytes_per_step": d2h / a.steps,
                    "note": "through fb_fillgaps_main: files -> model -> per-gap control on host threads -> engine; reference-equivalent pass-1 placements / wall",
                    "gaps_per_s": (len(open(os.path.join(case, "partial", "Temp", "gapInfo.txt")).readlines()) * world * a.steps) / dt_max,
                    "host_seconds_per_step": {k: sum(m[k] for m in metrics) / a.steps for k in ("t_load", "t_model", "t_prepare", "t_fill", "t_write", "t_ctx_upload", "t_workers", "t_engine_calls", "cpu_workers")}},
            "gpu_launches": int(launches_all), "device_placements_p1": dev_p1_all, "device_placements_p2": dev_p2_all, "roofline": roof}
    if world == 1 and not a.no_cpu_baseline and fc.have_reference():
        sample = prepare_case(os.path.join(base, "sample"), SAMPLE, 1102)
        placements = ref_equivalent_placements(sample, os.path.join(base, "sample_work"))
        secs = run_step_reference(sample, cores)
        # beside the as-shipped driver run (which sleeps 1 s after every worker start, FillGaps.cpp:675): the same workers
        # started together (steady state), and the tuned -O2 flavour
        steady = sum(fc.run_reference_workers_parallel(sample, mode, cores, "figbird_worker_O0") for mode in ("partial", "unmapped"))
        tuned = sum(fc.run_reference_workers_parallel(sample, mode, cores, "figbird_worker_O2") for mode in ("partial", "unmapped"))
        line["cpu_baseline"] = {"value": placements / secs, "unit": unit, "cores": cores, "kind": "reference",
                                "sample": "32-gap / 294 kbp sample of the c2 workload (same parameters), FillGaps partial+unmapped, as-shipped -O0 worker, numthreads=%d, %.1f s" % (cores, secs),
                                "steady_state": {"value": placements / steady, "seconds": steady, "note": "same sample, %d as-shipped -O0 worker processes started together (no driver sleep, no run-time compile)" % min(cores, 32)},
                                "tuned": {"value": placements / tuned, "seconds": tuned, "note": "same, worker built -O2 -D_FORTIFY_SOURCE=0 (plain -O2 aborts, SURVEY 5)"}}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
