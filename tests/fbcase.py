"""Test / bench harness: synthetic cases, the reference arm (oracle/_ref) and our arm.

A *case* is a directory produced by :func:`make_case`::

    case/draft.fa result1.sam result2.sam truth_gaps.txt        (tools/fbgen.cpp)
    case/partial/{myout.sam,Gaps/,Temp/}    inputs of FillGaps in partial mode  (reference Preprocess, mode 1)
    case/unmapped/{myout.sam,Gaps/,Temp/}   inputs of FillGaps in unmapped mode (reference Preprocess, mode 2)

Both arms then get the same argv as RunFigbird.sh:352 would pass.  Only tests/, bench.py and
__graft_entry__.py import this module; nothing in the product does.
"""
import os
import shutil
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
OBUILD = os.path.join(ROOT, "oracle", "_build")
TBUILD = os.path.join(ROOT, "tools", "_build")
PBUILD = os.path.join(ROOT, "figbird_b200", "_build")

OUTPUT_FILES = ("gapout.txt", "filledContigs.fa", "Ncount.txt", "draw.txt")


def sh(cmd, cwd=None, env=None, check=True, timeout=None):
    p = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    if check and p.returncode != 0:
        raise RuntimeError("command failed (%d): %s\n%s" % (p.returncode, " ".join(map(str, cmd)), p.stdout.decode(errors="replace")[-4000:]))
    return p


def build_tools():
    os.makedirs(TBUILD, exist_ok=True)
    exe = os.path.join(TBUILD, "fbgen")
    src = os.path.join(ROOT, "tools", "fbgen.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        sh(["g++", "-O2", "-std=c++17", "-w", "-pthread", "-o", exe, src])
    return exe


def build_oracle():
    """Builds oracle/_ref (when /root/reference exists) and oracle/_build (always)."""
    sh(["make", "-C", os.path.join(ROOT, "oracle"), "all"])


def have_reference():
    return os.path.exists(os.path.join(REF, "FillGaps")) and os.path.exists(os.path.join(REF, "Preprocess"))


def ref_env(worker="figbird_worker_O2"):
    env = dict(os.environ)
    env["PATH"] = os.path.join(REF, "shim") + os.pathsep + env.get("PATH", "")
    env["FB_WORKER"] = os.path.join(REF, worker)
    return env


def make_case(case, gen_args, readlen=100, insert=200, maxdist_local=None):
    """fbgen + reference Preprocess (modes 1 and 2).  gen_args: dict of fbgen options (without leading --)."""
    fbgen = build_tools()
    os.makedirs(case, exist_ok=True)
    args = [fbgen, "--out", case, "--readlen", str(readlen), "--insert", str(insert)]
    for k, v in gen_args.items():
        args += ["--" + k, str(v)]
    sh(args)
    x1 = int(gen_args.get("x1", insert)) if maxdist_local is None else maxdist_local
    x2 = int(1.15 * insert)
    draft = os.path.join(case, "draft.fa")
    for mode in ("partial", "unmapped"):
        d = os.path.join(case, mode)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(os.path.join(d, "Temp"))
    p = os.path.join(case, "partial")
    u = os.path.join(case, "unmapped")
    os.makedirs(os.path.join(p, "Gaps")); os.makedirs(os.path.join(u, "Gaps"))
    # the two Preprocess passes (RunFigbird.sh:285 mode 1, :338 mode 2) are independent: run them side by side
    cmds = [[os.path.join(REF, "Preprocess"), draft, str(x1), "1", os.path.join(case, "result1.sam"), os.path.join(p, "myout.sam"), draft,
             "r1.fq", "r2.fq", os.path.join(p, "Gaps") + "/", os.path.join(p, "Temp") + "/", "1", "0", "0"],
            [os.path.join(REF, "Preprocess"), draft, str(x2), "2", os.path.join(case, "result2.sam"), os.path.join(u, "myout.sam"), draft,
             "r1.fq", "r2.fq", os.path.join(u, "Gaps") + "/", os.path.join(u, "Temp") + "/", "1", "0", "0"]]
    procs = [subprocess.Popen(c, cwd=case, stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for c in cmds]
    for c, pr in zip(cmds, procs):
        out = pr.communicate()[0]
        if pr.returncode != 0:
            raise RuntimeError("command failed (%d): %s\n%s" % (pr.returncode, " ".join(c), out.decode(errors="replace")[-4000:]))
    # FillGaps needs partial_gaps_<g>.sam for every gap in unmapped mode too (Figbird.cpp:1801,1915): the script keeps one Gaps/
    # directory across both passes; here the unmapped run's directory gets links to the partial pass's files
    for f in os.listdir(os.path.join(p, "Gaps")):
        dst = os.path.join(u, "Gaps", f)
        if not os.path.exists(dst):
            os.link(os.path.join(p, "Gaps", f), dst)
    with open(os.path.join(case, "params.txt"), "w") as f:
        f.write("readlen %d\ninsert %d\nx1 %d\nx2 %d\n" % (readlen, insert, x1, x2))
    return case


def case_params(case):
    d = {}
    with open(os.path.join(case, "params.txt")) as f:
        for line in f:
            k, v = line.split()
            d[k] = int(v)
    return d


def fillgaps_argv(case, mode, tmp, threads=1, script_itr=1, neg_overlap=30):
    """argv[1..15] of FillGaps as RunFigbird.sh:352 builds it."""
    pr = case_params(case)
    d = os.path.join(case, mode)
    maxdist = pr["x1"] if mode == "partial" else pr["x2"]
    pf, um = ("1", "0") if mode == "partial" else ("0", "1")
    return [os.path.join(case, "draft.fa"), str(maxdist), str(pr["readlen"]), str(script_itr), pf, um, str(threads),
            os.path.join(d, "myout.sam"), tmp.rstrip("/") + "/", os.path.join(d, "Gaps") + "/", str(neg_overlap), str(pr["readlen"]), "0", "0", str(pr["insert"])]


def _fresh_tmp(case, mode, name):
    run = os.path.join(case, "run_" + name + "_" + mode)
    shutil.rmtree(run, ignore_errors=True)
    os.makedirs(os.path.join(run, "Temp"))
    for f in ("gapInfo.txt", "stat.txt", "stat2.txt"):
        shutil.copy(os.path.join(case, mode, "Temp", f), os.path.join(run, "Temp", f))
    return run


def read_outputs(tmp):
    out = {}
    for f in OUTPUT_FILES:
        p = os.path.join(tmp, f)
        out[f] = open(p, "rb").read() if os.path.exists(p) else None
    return out


def run_reference(case, mode, threads=1, worker="figbird_worker_O2", extra_env=None):
    """The unmodified reference FillGaps driver + prebuilt worker (through the g++ shim)."""
    run = _fresh_tmp(case, mode, "ref")
    env = ref_env(worker)
    if extra_env:
        env.update(extra_env)
    t0 = time.time()
    sh([os.path.join(REF, "FillGaps")] + fillgaps_argv(case, mode, os.path.join(run, "Temp"), threads), cwd=run, env=env)
    dt = time.time() - t0
    out = read_outputs(os.path.join(run, "Temp"))
    out["seconds"] = dt
    out["dir"] = run
    return out


def run_reference_worker(case, mode, worker="figbird_worker_dump", extra_env=None):
    """One reference worker over all gaps (Figbird.cpp:6957-6973 CLI), e.g. the dump-instrumented flavour."""
    run = _fresh_tmp(case, mode, "refw")
    tmp = os.path.join(run, "Temp")
    ngaps = sum(1 for _ in open(os.path.join(tmp, "gapInfo.txt")))
    with open(os.path.join(tmp, "gaploads.txt"), "w") as f:
        f.write("".join("%d\t" % g for g in range(ngaps)) + "\n")
    a = fillgaps_argv(case, mode, tmp)
    argv = [os.path.join(REF, worker), a[0], a[1], a[2], a[3], a[4], a[5], "0", str(ngaps), a[7], a[8], a[9], a[10], a[11], "400", a[13], a[14]]
    env = dict(os.environ)
    if extra_env:
        env.update(extra_env)
    t0 = time.time()
    sh(argv, cwd=run, env=env)
    return {"seconds": time.time() - t0, "dir": run, "gapout": open(os.path.join(tmp, "gapout0.txt"), "rb").read()}


def run_reference_workers_parallel(case, mode, threads, worker="figbird_worker_O0"):
    """Steady-state flavour of the reference: `threads` worker processes (Figbird.cpp:6957-6973 CLI) started
    together on a round-robin split of the gaps -- what FillGaps.cpp:668-679 does minus the run-time compile and the 1 s sleep
    after every thread start.  Returns wall seconds."""
    run = _fresh_tmp(case, mode, "refp")
    tmp = os.path.join(run, "Temp")
    ngaps = sum(1 for _ in open(os.path.join(tmp, "gapInfo.txt")))
    threads = max(1, min(threads, ngaps))
    loads = [[g for g in range(ngaps) if g % threads == t] for t in range(threads)]
    with open(os.path.join(tmp, "gaploads.txt"), "w") as f:
        for l in loads:
            f.write("".join("%d\t" % g for g in l) + "\n")
    a = fillgaps_argv(case, mode, tmp)
    t0 = time.time()
    procs = []
    for t in range(threads):
        argv = [os.path.join(REF, worker), a[0], a[1], a[2], a[3], a[4], a[5], str(t), str(len(loads[t])), a[7], a[8], a[9], a[10], a[11], "400", a[13], a[14]]
        procs.append(subprocess.Popen(argv, cwd=run, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    rc = [p.wait() for p in procs]
    dt = time.time() - t0
    if any(rc):
        raise RuntimeError("reference worker failed: %r" % rc)
    return dt


def run_ours(case, mode, exe, threads=1, extra_env=None, name="ours"):
    run = _fresh_tmp(case, mode, name)
    env = dict(os.environ)
    if os.path.dirname(os.path.abspath(exe)) == OBUILD:
        # CPU engines: no extra speculation at the tail of a run (it costs nothing on a GPU, minutes on the CPU restatement)
        env.setdefault("FIGBIRD_TAIL_ITEMS", "0")
    if extra_env:
        env.update(extra_env)
    t0 = time.time()
    p = sh([exe] + fillgaps_argv(case, mode, os.path.join(run, "Temp"), threads), cwd=run, env=env)
    dt = time.time() - t0
    out = read_outputs(os.path.join(run, "Temp"))
    out["seconds"] = dt
    out["dir"] = run
    out["log"] = p.stdout.decode(errors="replace")
    return out


def oracle_exe():
    return os.path.join(OBUILD, "fillgaps_oracle")


def product_exe():
    return os.path.join(PBUILD, "fillgaps")


def draw_by_gap(data):
    """draw.txt split per gap (the reference concatenates it in worker order)."""
    out, cur, key = {}, [], None
    for line in (data or b"").split(b"\n"):
        if b"+Gap = " in line:
            if key is not None:
                out[key] = b"\n".join(cur)
            key = int(line.split(b"+Gap = ")[1].split(b" ")[0])
            cur = []
        cur.append(line)
    if key is not None:
        out[key] = b"\n".join(cur)
    # (the text of a gap ends with a newline; where it ends, relative to the next gap's header or the end of the file, differs with the order)
    return {k: v.rstrip(b"\n") for k, v in out.items()}
