"""CPU: bench.py's bookkeeping that needs no GPU -- the JSON line assembled from recorded per-call metrics (contract keys, the
breakdown, the at-size parity sample), the sample specs of one- and two-library workloads, the cached sample counts."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def recorded_metrics(ngpu):
    # what FIGBIRD_METRICS holds after a partial + unmapped call pair (fb_fillgaps.cpp), summed by run_step_ours; N = 4 run of C4
    per = [7856.4, 7804.8, 7783.9, 7772.1][:ngpu]
    return {"engine": "cuda-sm100a", "gaps": 20000, "gpus": 3 * ngpu, "t_load": 0.107, "t_model": 0.447, "t_prepare": 0.096, "t_fill": 8.371, "t_write": 0.162,
            "t_engine_calls": 7.729, "t_result_copy": 0.0, "t_ctx_upload": 0.246, "t_workers": 8.120, "cpu_workers": 15.19,
            "ref_placements_p1": 189881049860, "dev_placements_p1": 194016450946, "dev_placements_p2": 194128058866, "dev_base_terms": 58157580190002,
            "kernel_launches": 2100, "device_ms": max(per), "h2d_bytes": 70000000, "d2h_bytes": 11911434384, "ticks": 640,
            "lane_steps_p1": 22803361370368, "lane_steps_p2": 484659557632, "device_ms_per_gpu": per}


def test_line_from_recorded_metrics(tmp_path):
    a = argparse.Namespace(steps=2, warmup=1, workload="c4", gpus=4)
    exp = json.load(open(os.path.join(ROOT, "tests", "golden", "c4_sample.json")))
    work = str(tmp_path / "work")
    for mode in ("partial", "unmapped"):
        os.makedirs(os.path.join(work, mode, "Temp"))
        lines = dict(exp[mode])
        if mode == "unmapped":
            g = sorted(lines)[0]; lines[g] = lines[g].replace("\t", "\tX", 1)      # one sampled gap differs
        with open(os.path.join(work, mode, "Temp", "gapout.txt"), "w") as f:
            f.write("".join(lines.values()))
    clk = {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 40}
    mb = {"dmul_tinstr_s": 17.83, "dfma_tflops": 34.7}
    line = bench.assemble_line(a, 4, [recorded_metrics(4), recorded_metrics(4)], 17.84, clk, mb, 148, 10000, work, "placements/s", "m", {"workload": "c4"}, 31.0)
    json.dumps(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in line, k
    assert line["n_gpus"] == 4 and line["scaling"] == "strong" and line["gpu_launches"] == 4200
    assert abs(line["ms_per_step"] - 8920.0) < 1e-6 and abs(line["e2e"]["gaps_per_s"] - 20000 / 17.84) < 1e-9
    assert abs(line["value"] - 2 * 189881049860 / (2 * 7.8564)) / line["value"] < 1e-12      # reference-equivalent placements / kernel time of the busiest GPU
    r = line["roofline"]
    assert 0.2 < r["frac"] < 0.3 and r["unit"] == "TFLOP/s" and r["smem_pipe"]["frac"] > 0.5 and r["traffic"] > 1e6
    b = line["e2e"]["seconds_per_step_breakdown"]
    assert abs(sum(b.values()) - 8.92) < 1e-6 and line["e2e"]["limiter"].startswith("kernels")
    assert len(line["e2e"]["device_ms_per_gpu_per_step"]) == 4
    ps = line["parity_sample"]
    assert ps["gaps_sampled"] == 120 and ps["identical_partial"] == 120 and ps["identical_unmapped"] == 119 and len(ps["different_unmapped"]) == 1


def test_sample_specs_and_cached_counts():
    one = bench.sample_spec("c4", 16)
    assert one[0]["gaps"] == 16 and one[0]["genome"] == 160000 and one[1:] == (150, 500) and "threads" not in one[0]
    two = bench.sample_spec("c3", 16)
    assert set(two) == {"partial", "unmapped"} and two["unmapped"][2] == 3500 and two["partial"][0]["gaps"] == 16
    assert bench.gen_of("c5s")["gaps"] == 600
    key = "c4_sample%d_seed%d" % (bench.SAMPLE_GAPS["c4"][1], bench.SEEDS["c4"] + 2000)
    assert bench.count_reference_placements("/nonexistent", 1, key) == json.load(open(os.path.join(ROOT, "tests", "golden", "bench_sample_counts.json")))[key]
