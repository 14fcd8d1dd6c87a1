"""ctypes binding of oracle/liboracle.so -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  The product package (addapt_b200/) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_PAR = os.path.join(os.path.dirname(HERE), "params", "rna_turner2004_reconstructed.par")
ORC_INF = 10000000


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("fold_oracle.c", "design_oracle.c", "batch_oracle.c", "fold_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s", "liboracle.so"])
    return so


class _Motif(C.Structure):
    _fields_ = [("seq", C.c_char_p), ("fold", C.c_char_p), ("energy_kcal", C.c_double)]


class _BpTerm(C.Structure):
    _fields_ = [("kind", C.c_int), ("cond", C.c_int), ("favorable", C.c_int), ("weight", C.c_double), ("i", C.c_int), ("j", C.c_int),
                ("n_pos", C.c_int), ("pos", C.POINTER(C.c_int))]


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("init_seq", C.c_char_p), ("n_macro", C.c_int), ("macro_cst", C.POINTER(C.c_char_p)),
                ("n_terms", C.c_int), ("term_macro", C.POINTER(C.c_int)), ("term_cond", C.POINTER(C.c_int)),
                ("term_favorable", C.POINTER(C.c_int)), ("term_weight", C.POINTER(C.c_double)), ("motif", _Motif),
                ("n_ctx", C.c_int), ("ctx_before", C.POINTER(C.c_char_p)), ("ctx_after", C.POINTER(C.c_char_p)),
                ("n_bp_terms", C.c_int), ("bp_terms", C.POINTER(_BpTerm))]


def _b(s):
    return s if isinstance(s, bytes) else s.encode()


class Oracle:
    def __init__(self, par_path=None):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_model_load.restype = C.c_void_p
        L.orc_model_load.argtypes = [C.c_char_p]
        L.orc_error.restype = C.c_char_p
        L.orc_kT_kcal.restype = C.c_double
        L.orc_kT_kcal.argtypes = [C.c_void_p]
        L.orc_eval_structure.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p]
        L.orc_motif_contribution.argtypes = [C.c_void_p, C.POINTER(_Motif)]
        L.orc_mfe.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(_Motif), C.POINTER(C.c_int32), C.c_char_p]
        L.orc_pf.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(_Motif), C.POINTER(C.c_double), C.c_void_p]
        L.orc_macrostate_prob.restype = C.c_double
        L.orc_macrostate_prob.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(_Motif), C.c_int]
        L.orc_count_terms.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_score.restype = C.c_double
        L.orc_score.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_char_p, C.c_int, C.c_void_p]
        L.orc_philox4x32.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_mc_run.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double),
                                 C.c_int, C.c_char_p, C.POINTER(C.c_double), C.c_char_p, C.POINTER(C.c_double),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_void_p]
        L.orc_mc_run_trace.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double),
                                       C.c_int, C.c_char_p, C.POINTER(C.c_double), C.c_char_p, C.POINTER(C.c_double),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_int, C.c_void_p, C.c_void_p]
        L.orc_mutate.argtypes = [C.POINTER(_Problem), C.c_char_p, C.c_int, C.c_char]
        L.orc_pf_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(_Motif), C.c_void_p, C.c_void_p, C.c_int]
        L.orc_mfe_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(_Motif), C.c_void_p, C.c_void_p, C.c_int]
        L.orc_score_batch.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        # same precedence as the engine: explicit argument > $ADDAPT_B200_PAR > the reconstructed file in params/
        self.par_path = par_path or os.environ.get("ADDAPT_B200_PAR") or DEFAULT_PAR
        with open(self.par_path) as f:
            self.param_provenance = "reconstructed" if "RECONSTRUCTION" in "".join(f.readlines()[:12]) else "file"
        self.model = L.orc_model_load(_b(self.par_path))
        if not self.model:
            raise RuntimeError("oracle: " + L.orc_error().decode())
        self.kT = L.orc_kT_kcal(self.model)

    # ---- helpers -------------------------------------------------------------------------
    def _motif(self, motif):
        if motif is None:
            return None
        seq, fold, kd_uM = motif
        import math
        m = _Motif(_b(seq), _b(fold), self.kT * math.log(kd_uM / 1e6))
        return C.byref(m), m

    def num_threads(self):
        return self.lib.orc_num_threads()

    def eval_structure(self, seq, structure):
        return self.lib.orc_eval_structure(self.model, len(seq), _b(seq), _b(structure))

    def motif_contribution(self, motif):
        ref, keep = self._motif(motif)
        return self.lib.orc_motif_contribution(self.model, ref)

    def mfe(self, seq, cst=None, motif=None):
        n = len(seq)
        e = C.c_int32()
        st = C.create_string_buffer(n + 1)
        mo = self._motif(motif)
        rc = self.lib.orc_mfe(self.model, n, _b(seq), _b(cst) if cst else None, mo[0] if mo else None, C.byref(e), st)
        if rc:
            raise ValueError(self.lib.orc_error().decode())
        return e.value, st.value.decode()

    def pf(self, seq, cst=None, motif=None, bppm=False):
        n = len(seq)
        dg = C.c_double()
        bp = np.zeros(n * (n + 1) // 2, dtype=np.float64) if bppm else None
        mo = self._motif(motif)
        rc = self.lib.orc_pf(self.model, n, _b(seq), _b(cst) if cst else None, mo[0] if mo else None, C.byref(dg),
                             bp.ctypes.data if bppm else None)
        if rc:
            raise ValueError(self.lib.orc_error().decode())
        return (dg.value, bp) if bppm else dg.value

    def macrostate_prob(self, seq, cst, motif=None, float_dg=False):
        mo = self._motif(motif)
        return self.lib.orc_macrostate_prob(self.model, len(seq), _b(seq), _b(cst), mo[0] if mo else None, int(float_dg))

    def count_terms(self, seq):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_count_terms(len(seq), _b(seq), C.byref(a), C.byref(b))
        return a.value, b.value

    def philox(self, key0, key1, ctr):
        c = (C.c_uint32 * 4)(*ctr)
        o = (C.c_uint32 * 4)()
        self.lib.orc_philox4x32(key0, key1, c, o)
        return list(o)

    # ---- design problem ------------------------------------------------------------------
    def problem(self, prob):
        """prob: dict(seq, macrostates={name: cst}, terms=[(macro_name, cond 'apo'|'holo', favorable bool, weight)],
        motif=(seq, fold, kd_uM)|None, contexts=[(before, after)])."""
        import math
        names = list(prob["macrostates"].keys())
        keep = {}
        keep["macro"] = (C.c_char_p * len(names))(*[_b(prob["macrostates"][k]) for k in names])
        T = prob["terms"]
        keep["tm"] = (C.c_int * len(T))(*[names.index(t[0]) for t in T])
        keep["tc"] = (C.c_int * len(T))(*[1 if t[1] == "holo" else 0 for t in T])
        keep["tf"] = (C.c_int * len(T))(*[1 if t[2] else 0 for t in T])
        keep["tw"] = (C.c_double * len(T))(*[float(t[3]) for t in T])
        ctx = prob.get("contexts") or []
        keep["cb"] = (C.c_char_p * max(len(ctx), 1))(*([_b(c[0]) for c in ctx] or [b""]))
        keep["ca"] = (C.c_char_p * max(len(ctx), 1))(*([_b(c[1]) for c in ctx] or [b""]))
        mo = prob.get("motif")
        m = _Motif(_b(mo[0]), _b(mo[1]), self.kT * math.log(mo[2] / 1e6)) if mo else _Motif(None, None, 0.0)
        bts = prob.get("bp_terms") or []
        bp = (_BpTerm * max(len(bts), 1))()
        keep["bp"], keep["bp_pos"] = bp, []
        for q, t in enumerate(bts):      # ("pair", cond, favorable, weight, i, j) | ("density", cond, favorable, weight, [positions])
            bp[q].kind, bp[q].cond, bp[q].favorable, bp[q].weight = (0 if t[0] == "pair" else 1), (1 if t[1] == "holo" else 0), int(bool(t[2])), float(t[3])
            if t[0] == "pair":
                bp[q].i, bp[q].j, bp[q].n_pos, bp[q].pos = int(t[4]), int(t[5]), 0, None
            else:
                arr = (C.c_int * len(t[4]))(*[int(x) for x in t[4]])
                keep["bp_pos"].append(arr)
                bp[q].i, bp[q].j, bp[q].n_pos, bp[q].pos = 0, 0, len(t[4]), arr
        P = _Problem(len(prob["seq"]), _b(prob["seq"]), len(names), keep["macro"], len(T), keep["tm"], keep["tc"], keep["tf"],
                     keep["tw"], m, len(ctx), keep["cb"], keep["ca"], len(bts), bp)
        keep["P"] = P
        return P, keep

    def score(self, prob, seq, float_dg=False, want_terms=False):
        P, keep = self.problem(prob)
        nt = (P.n_terms + P.n_bp_terms) * max(P.n_ctx, 1)
        terms = np.zeros(nt)
        s = self.lib.orc_score(self.model, C.byref(P), _b(seq), int(float_dg), terms.ctypes.data)
        return (s, terms) if want_terms else s

    def score_batch(self, prob, seqs, float_dg=False, threads=0):
        P, keep = self.problem(prob)
        B = len(seqs)
        flat = b"".join(_b(s) for s in seqs)
        out = np.zeros(B)
        terms = np.zeros((B, (P.n_terms + P.n_bp_terms) * max(P.n_ctx, 1)))
        self.lib.orc_score_batch(self.model, C.byref(P), B, flat, int(float_dg), out.ctypes.data, terms.ctypes.data, threads)
        return out, terms

    def mutate(self, prob, seq, pos, nucleotide):
        """≙ mutate_recursively (reference src/sampling.cc:205-282); raises ValueError where the reference throws"""
        P, keep = self.problem(prob)
        buf = C.create_string_buffer(_b(seq), len(seq) + 1)
        rc = self.lib.orc_mutate(C.byref(P), buf, pos, _b(nucleotide))
        if rc:
            raise ValueError("mutate_recursively failed rc=%d" % rc)
        return buf.value.decode()

    def mc_run(self, prob, seed, traj_id, n_steps, thermostat=("annealing", 5.0, 0.0, 300), float_dg=False, trace=False, trace_interval=0):
        P, keep = self.problem(prob)
        n = P.n
        if thermostat[0] == "fixed":
            tk, ta = 0, (C.c_double * 3)(thermostat[1], 0, 0)
        elif thermostat[0] == "auto":            # (target acceptance rate, training period, initial temperature)
            tk, ta = 2, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        else:
            tk, ta = 1, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        fs, bs = C.create_string_buffer(n + 1), C.create_string_buffer(n + 1)
        fsc, bsc, bst = C.c_double(), C.c_double(), C.c_int32()
        oc = (C.c_uint32 * 4)()
        tr = np.zeros(n_steps) if trace else None
        tv = tq = None
        if trace_interval > 0 and n_steps > 0:
            cap, width = (n_steps + trace_interval - 1) // trace_interval, 8 + (P.n_terms + P.n_bp_terms) * max(P.n_ctx, 1)
            tv, tq = np.zeros((cap, width)), np.zeros((cap, 2, n), dtype=np.uint8)
            rc = self.lib.orc_mc_run_trace(self.model, C.byref(P), seed, traj_id, n_steps, tk, ta, int(float_dg), fs, C.byref(fsc), bs,
                                           C.byref(bsc), C.byref(bst), oc, trace_interval, tv.ctypes.data, tq.ctypes.data)
        else:
            rc = self.lib.orc_mc_run(self.model, C.byref(P), seed, traj_id, n_steps, tk, ta, int(float_dg), fs, C.byref(fsc), bs,
                                     C.byref(bsc), C.byref(bst), oc, tr.ctypes.data if trace else None)
        if rc:
            raise ValueError("mc_run failed rc=%d" % rc)
        return dict(final_seq=fs.value.decode(), final_score=fsc.value, best_seq=bs.value.decode(), best_score=bsc.value,
                    best_step=bst.value, outcomes=list(oc), trace=tr, trace_values=tv, trace_seqs=tq)

    # ---- batches (threaded) --------------------------------------------------------------
    def pf_batch(self, seqs, cst=None, motif=None, bppm=False, threads=0):
        B, n = len(seqs), len(seqs[0])
        flat = b"".join(_b(s) for s in seqs)
        dg = np.zeros(B)
        bp = np.zeros((B, n * (n + 1) // 2)) if bppm else None
        mo = self._motif(motif)
        rc = self.lib.orc_pf_batch(self.model, B, n, flat, _b(cst) if cst else None, mo[0] if mo else None, dg.ctypes.data,
                                   bp.ctypes.data if bppm else None, threads)
        if rc:
            raise ValueError(self.lib.orc_error().decode())
        return (dg, bp) if bppm else dg

    def mfe_batch(self, seqs, cst=None, motif=None, threads=0):
        B, n = len(seqs), len(seqs[0])
        flat = b"".join(_b(s) for s in seqs)
        e = np.zeros(B, dtype=np.int32)
        st = np.zeros((B, n), dtype=np.uint8)
        mo = self._motif(motif)
        rc = self.lib.orc_mfe_batch(self.model, B, n, flat, _b(cst) if cst else None, mo[0] if mo else None, e.ctypes.data,
                                    st.ctypes.data, threads)
        if rc:
            raise ValueError(self.lib.orc_error().decode())
        return e, [bytes(r).decode() for r in st]


def tri(n, i, j):
    if i > j:
        i, j = j, i
    return i * n - i * (i - 1) // 2 + (j - i)
