"""ctypes binding of include/figbird_b200.h.

This is the stub a Python host would write against the C ABI; it adds nothing to it.  The product library
is ``figbird_b200/_build/libfigbird_b200.so`` (CUDA, sm_100a).  There is no fallback: if the library is
missing or no B200-class GPU is present, loading / ``Engine()`` raises.  Tests may pass ``lib_path`` to
load the oracle library that implements the same ABI on the CPU (``oracle/_build/libfb_oracle.so``).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "_build", "libfigbird_b200.so")

FB_MODE_PARTIAL, FB_MODE_UNMAPPED = 0, 1
FB_READ_LEFT, FB_READ_REVERSE, FB_READ_NOMATE = 1, 2, 4
FB_ITEM_EM, FB_ITEM_HARD = 0, 1
FB_FLAG_EXTRA_PASS, FB_FLAG_RECORD_ALL, FB_FLAG_WANT_COUNTS, FB_FLAG_RESUME, FB_FLAG_NO_COMP_STOP, FB_FLAG_FINALIZE_REF = 1, 2, 4, 8, 16, 32


class FbModel(C.Structure):
    _fields_ = [("max_read_len", C.c_int32), ("err_pos", C.POINTER(C.c_double)), ("ins_pos", C.POINTER(C.c_double)),
                ("del_pos", C.POINTER(C.c_double)), ("err_type", C.c_double * 25), ("n_insert", C.c_int32),
                ("insert_pdf", C.POINTER(C.c_double)), ("insert_min", C.c_int32), ("insert_max", C.c_int32), ("prob_cutoff", C.c_int32)]


class FbGap(C.Structure):
    _fields_ = [("gap_start", C.c_int64), ("mode", C.c_int32), ("orig_len", C.c_int32), ("n_reads", C.c_int32), ("read_begin", C.c_int32),
                ("flank_len", C.c_int32), ("flank_begin", C.c_int32), ("pile_len", C.c_int32), ("pile_begin", C.c_int32)]


class FbGapBatch(C.Structure):
    _fields_ = [("n_gaps", C.c_int32), ("gaps", C.POINTER(FbGap)), ("n_reads", C.c_int32), ("read_len", C.POINTER(C.c_int32)),
                ("read_code_off", C.POINTER(C.c_int64)), ("read_mate", C.POINTER(C.c_int32)), ("read_flags", C.POINTER(C.c_uint8)),
                ("read_jlo", C.POINTER(C.c_uint8)), ("read_jcut", C.POINTER(C.c_uint8)), ("n_codes", C.c_int64), ("read_codes", C.POINTER(C.c_uint8)),
                ("n_flank", C.c_int64), ("flank_codes", C.POINTER(C.c_uint8)), ("n_pile_rows", C.c_int64), ("pile_left", C.POINTER(C.c_int32)),
                ("pile_right", C.POINTER(C.c_int32))]


class FbWorkItem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("gap", C.c_int32), ("cand_len", C.c_int32), ("max_rounds", C.c_int32), ("flags", C.c_int32),
                ("comp_count_in", C.c_int32), ("counts_in", C.POINTER(C.c_double)), ("string_in", C.POINTER(C.c_uint8))]


class FbItemOut(C.Structure):
    _fields_ = [("calls", C.c_int32), ("comp_count", C.c_int32), ("flags", C.c_int32), ("n_reads", C.c_int32), ("cand_len", C.c_int32),
                ("n_slots", C.c_int32), ("placements", C.c_int64), ("off_p1max", C.c_int64), ("off_p2max", C.c_int64), ("off_pos2", C.c_int64),
                ("off_soft", C.c_int64), ("off_hard", C.c_int64), ("off_cov", C.c_int64), ("off_counts", C.c_int64)]


class FbCounters(C.Structure):
    _fields_ = [("placements_p1", C.c_int64), ("placements_p2", C.c_int64), ("base_terms", C.c_int64), ("kernel_launches", C.c_int64),
                ("device_ms", C.c_double), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("lane_steps_p1", C.c_int64), ("lane_steps_p2", C.c_int64), ("device_union_ms", C.c_double)]


EXPORTS = ["fb_ctx_create", "fb_ctx_destroy", "fb_ctx_set_latency_critical", "fb_last_error", "fb_engine_name", "fb_model_upload", "fb_batch_upload", "fb_em_run",
           "fb_get_counters", "fb_microbench_fp64", "fb_fillgaps_main",
           "fb_preprocess_main", "fb_combinegaps_main", "fb_flanktrim_main", "fb_reduce_scf_main", "fb_reverse_main"]


def load(lib_path=None):
    path = lib_path or PRODUCT_LIB
    if not os.path.exists(path):
        raise RuntimeError("figbird_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    lib.fb_ctx_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]; lib.fb_ctx_create.restype = C.c_int32
    lib.fb_ctx_destroy.argtypes = [C.c_void_p]; lib.fb_ctx_destroy.restype = None
    lib.fb_ctx_set_latency_critical.argtypes = [C.c_void_p, C.c_int32]; lib.fb_ctx_set_latency_critical.restype = C.c_int32
    lib.fb_last_error.argtypes = [C.c_void_p]; lib.fb_last_error.restype = C.c_char_p
    lib.fb_engine_name.argtypes = []; lib.fb_engine_name.restype = C.c_char_p
    lib.fb_model_upload.argtypes = [C.c_void_p, C.POINTER(FbModel)]; lib.fb_model_upload.restype = C.c_int32
    lib.fb_batch_upload.argtypes = [C.c_void_p, C.POINTER(FbGapBatch)]; lib.fb_batch_upload.restype = C.c_int32
    lib.fb_em_run.argtypes = [C.c_void_p, C.POINTER(FbWorkItem), C.c_int32, C.POINTER(C.POINTER(FbItemOut))]; lib.fb_em_run.restype = C.c_int32
    lib.fb_get_counters.argtypes = [C.c_void_p, C.POINTER(FbCounters)]; lib.fb_get_counters.restype = C.c_int32
    lib.fb_microbench_fp64.argtypes = [C.c_void_p, C.POINTER(C.c_double)]; lib.fb_microbench_fp64.restype = C.c_int32
    lib.fb_fillgaps_main.argtypes = [C.c_int32, C.POINTER(C.c_char_p)]; lib.fb_fillgaps_main.restype = C.c_int32
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Engine:
    """One engine context (one GPU).  Arrays are numpy; see include/figbird_b200.h for the meaning of each field."""

    def __init__(self, device=0, lib_path=None):
        self.lib = load(lib_path)
        h = C.c_void_p()
        st = self.lib.fb_ctx_create(device, C.byref(h))
        if st != 0 or not h:
            msg = self.lib.fb_last_error(h).decode() if h else "no usable device"
            if h:
                self.lib.fb_ctx_destroy(h)
            raise RuntimeError("fb_ctx_create failed (%d): %s" % (st, msg))
        self.h = h
        self._keep = []

    def name(self):
        return self.lib.fb_engine_name().decode()

    def close(self):
        if self.h:
            self.lib.fb_ctx_destroy(self.h)
            self.h = None

    def _check(self, st, what):
        if st != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, st, self.lib.fb_last_error(self.h).decode()))

    def upload_model(self, err_pos, ins_pos, del_pos, err_type, insert_pdf, insert_min, insert_max, prob_cutoff):
        e = np.ascontiguousarray(err_pos, dtype=np.float64); i = np.ascontiguousarray(ins_pos, dtype=np.float64)
        d = np.ascontiguousarray(del_pos, dtype=np.float64); pdf = np.ascontiguousarray(insert_pdf, dtype=np.float64)
        m = FbModel()
        m.max_read_len = len(e); m.err_pos = _p(e, C.c_double); m.ins_pos = _p(i, C.c_double); m.del_pos = _p(d, C.c_double)
        et = np.ascontiguousarray(err_type, dtype=np.float64).reshape(25)
        for k in range(25):
            m.err_type[k] = float(et[k])
        m.n_insert = len(pdf); m.insert_pdf = _p(pdf, C.c_double)
        m.insert_min, m.insert_max, m.prob_cutoff = int(insert_min), int(insert_max), int(prob_cutoff)
        self._check(self.lib.fb_model_upload(self.h, C.byref(m)), "fb_model_upload")

    def upload_batch(self, gaps, reads):
        """gaps: list of dicts(gap_start, mode, orig_len, flank(2F codes), pile_left[T,4], pile_right[T,4], reads=[indices]);
        reads: list of dicts(codes, mate, flags, jlo, jcut), already grouped per gap in order."""
        ga = (FbGap * len(gaps))()
        rlen, roff, rmate, rfl, rjlo, rjcut, codes, flank, pl, pr = [], [], [], [], [], [], [], [], [], []
        ncodes = 0
        for gi, g in enumerate(gaps):
            G = ga[gi]
            G.gap_start, G.mode, G.orig_len = int(g["gap_start"]), int(g["mode"]), int(g["orig_len"])
            G.n_reads, G.read_begin = len(g["reads"]), len(rlen)
            F = len(g["flank"]) // 2
            G.flank_len, G.flank_begin = F, sum(len(x) for x in flank)
            T = len(g["pile_left"])
            G.pile_len, G.pile_begin = T, sum(len(x) for x in pl)
            flank.append(np.asarray(g["flank"], dtype=np.uint8)); pl.append(np.asarray(g["pile_left"], dtype=np.int32).reshape(T, 4))
            pr.append(np.asarray(g["pile_right"], dtype=np.int32).reshape(T, 4))
            for r in g["reads"]:
                c = np.asarray(r["codes"], dtype=np.uint8)
                rlen.append(len(c)); roff.append(ncodes); ncodes += len(c); codes.append(c)
                rmate.append(int(r["mate"])); rfl.append(int(r["flags"])); rjlo.append(int(r.get("jlo", 0))); rjcut.append(int(r.get("jcut", 0)))
        def cat(xs, dt, empty):
            return np.ascontiguousarray(np.concatenate(xs) if xs else np.asarray(empty), dtype=dt)
        a_rlen = cat([np.asarray(rlen)], np.int32, [0]) if rlen else np.zeros(1, np.int32)
        a_roff = np.asarray(roff if roff else [0], dtype=np.int64); a_rmate = np.asarray(rmate if rmate else [0], dtype=np.int32)
        a_rfl = np.asarray(rfl if rfl else [0], dtype=np.uint8); a_jlo = np.asarray(rjlo if rjlo else [0], dtype=np.uint8); a_jcut = np.asarray(rjcut if rjcut else [0], dtype=np.uint8)
        a_codes = cat(codes, np.uint8, [4]); a_flank = cat(flank, np.uint8, [4])
        a_pl = cat([x.reshape(-1) for x in pl], np.int32, [0, 0, 0, 0]); a_pr = cat([x.reshape(-1) for x in pr], np.int32, [0, 0, 0, 0])
        b = FbGapBatch()
        b.n_gaps, b.gaps, b.n_reads = len(gaps), ga, len(rlen)
        b.read_len, b.read_code_off, b.read_mate = _p(a_rlen, C.c_int32), _p(a_roff, C.c_int64), _p(a_rmate, C.c_int32)
        b.read_flags, b.read_jlo, b.read_jcut = _p(a_rfl, C.c_uint8), _p(a_jlo, C.c_uint8), _p(a_jcut, C.c_uint8)
        b.n_codes, b.read_codes, b.n_flank, b.flank_codes = len(a_codes), _p(a_codes, C.c_uint8), len(a_flank), _p(a_flank, C.c_uint8)
        b.n_pile_rows, b.pile_left, b.pile_right = len(a_pl) // 4, _p(a_pl, C.c_int32), _p(a_pr, C.c_int32)
        self._check(self.lib.fb_batch_upload(self.h, C.byref(b)), "fb_batch_upload")

    def run(self, items):
        """items: list of dicts(kind, gap, cand_len, max_rounds, flags, comp_count_in, counts_in, string_in) -> list of result dicts."""
        n = len(items)
        wi = (FbWorkItem * n)()
        keep = []
        for i, it in enumerate(items):
            w = wi[i]
            w.kind, w.gap, w.cand_len = int(it.get("kind", FB_ITEM_EM)), int(it["gap"]), int(it["cand_len"])
            w.max_rounds, w.flags, w.comp_count_in = int(it.get("max_rounds", 0)), int(it.get("flags", 0)), int(it.get("comp_count_in", 0))
            if it.get("counts_in") is not None:
                a = np.ascontiguousarray(it["counts_in"], dtype=np.float64); keep.append(a); w.counts_in = _p(a, C.c_double)
            if it.get("string_in") is not None:
                a = np.ascontiguousarray(it["string_in"], dtype=np.uint8); keep.append(a); w.string_in = _p(a, C.c_uint8)
        outs = (C.POINTER(FbItemOut) * n)()
        self._check(self.lib.fb_em_run(self.h, wi, n, outs), "fb_em_run")
        res = []
        for i in range(n):
            H = outs[i].contents
            base = C.addressof(H)
            R, Lg, S = H.n_reads, H.cand_len, H.n_slots
            def arr(off, dt, cnt):
                if cnt == 0:
                    return np.zeros(0, dtype=dt)
                buf = (C.c_char * (np.dtype(dt).itemsize * cnt)).from_address(base + off)
                return np.frombuffer(buf, dtype=dt, count=cnt).copy()
            r = dict(calls=H.calls, comp_count=H.comp_count, flags=H.flags, n_reads=R, cand_len=Lg, n_slots=S, placements=H.placements,
                     p1max=arr(H.off_p1max, np.float64, S * R).reshape(S, R), p2max=arr(H.off_p2max, np.float64, S * R).reshape(S, R),
                     pos2=arr(H.off_pos2, np.int32, S * R).reshape(S, R), soft=arr(H.off_soft, np.uint8, Lg), hard=arr(H.off_hard, np.uint8, Lg),
                     cov=arr(H.off_cov, np.int32, Lg), counts=arr(H.off_counts, np.float64, 5 * Lg).reshape(Lg, 5) if H.off_counts >= 0 else None)
            res.append(r)
        return res

    def microbench_fp64(self):
        out = (C.c_double * 2)()
        self._check(self.lib.fb_microbench_fp64(self.h, out), "fb_microbench_fp64")
        return {"dmul_tinstr_s": out[0], "dfma_tflops": out[1]}

    def counters(self):
        c = FbCounters()
        self._check(self.lib.fb_get_counters(self.h, C.byref(c)), "fb_get_counters")
        return {k: getattr(c, k) for k, _ in FbCounters._fields_}


def fillgaps(argv, lib_path=None):
    """fb_fillgaps_main in-process: argv = the 15 FillGaps arguments (without argv[0])."""
    lib = load(lib_path)
    full = [b"fillgaps"] + [a.encode() if isinstance(a, str) else a for a in argv]
    arr = (C.c_char_p * len(full))(*full)
    return lib.fb_fillgaps_main(len(full), arr)


def tool(name, argv, lib_path=None):
    """The host-only pipeline tools in-process: name in preprocess / combinegaps / flanktrim / reduce_scf / reverse,
    argv = that reference program's positional arguments (without argv[0]).  Returns its exit status."""
    lib = load(lib_path)
    fn = getattr(lib, "fb_%s_main" % name)
    fn.argtypes = [C.c_int32, C.POINTER(C.c_char_p)]; fn.restype = C.c_int32
    full = [name.encode()] + [a.encode() if isinstance(a, str) else a for a in argv]
    arr = (C.c_char_p * len(full))(*full)
    return fn(len(full), arr)
