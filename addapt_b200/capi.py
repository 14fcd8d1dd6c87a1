"""ctypes binding of libaddapt_b200.so (include/addapt_b200.h).  Plumbing only -- no arithmetic here.

Arrays may be numpy arrays (host) or anything exposing ``data_ptr()`` (torch CUDA tensors): the engine
detects host vs device pointers itself.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEFAULT_PAR = os.path.join(ROOT, "params", "rna_turner2004_reconstructed.par")

# the theophylline aptamer of the reference's tests (tests/test_scoring.cc:45-48)
THEOPHYLLINE = ("GAUACCAGCCGAAAGGCCCUUGGCAGC", "(...((.(((....)))....))...)", 0.32)


class EngineError(RuntimeError):
    pass


class _Motif(C.Structure):
    _fields_ = [("seq", C.c_char_p), ("fold", C.c_char_p), ("kd_uM", C.c_double)]


class _BpTerm(C.Structure):
    _fields_ = [("kind", C.c_int), ("cond", C.c_int), ("favorable", C.c_int), ("weight", C.c_double), ("i", C.c_int), ("j", C.c_int),
                ("n_pos", C.c_int), ("pos", C.POINTER(C.c_int))]


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("init_seq", C.c_char_p), ("n_macro", C.c_int), ("macro_cst", C.POINTER(C.c_char_p)),
                ("n_terms", C.c_int), ("term_macro", C.POINTER(C.c_int)), ("term_cond", C.POINTER(C.c_int)),
                ("term_favorable", C.POINTER(C.c_int)), ("term_weight", C.POINTER(C.c_double)), ("motif", _Motif),
                ("n_ctx", C.c_int), ("ctx_before", C.POINTER(C.c_char_p)), ("ctx_after", C.POINTER(C.c_char_p)),
                ("float_dg", C.c_int), ("n_bp_terms", C.c_int), ("bp_terms", C.POINTER(_BpTerm))]


def _b(s):
    return s if isinstance(s, bytes) else s.encode()


class Problem:
    """A design problem: ≙ Device + ScoreFunction of the reference (include/model.hh, include/scoring.hh).

    seq: mixed-case sequence (upper = mutable); macrostates: {name: constraint}; terms: list of
    (macrostate name, 'apo'|'holo', favorable bool, weight); motif: (seq, fold, kd_uM) or None;
    contexts: list of (before, after); bp_terms: base-pair-probability terms, each ("pair", 'apo'|'holo', favorable, weight, i, j)
    or ("density", 'apo'|'holo', favorable, weight, [positions]) with 0-based positions of the raw device (include/addapt_b200.h).
    Term values come back as [context][macrostate terms..., bp terms...].
    """

    def __init__(self, seq, macrostates, terms, motif=None, contexts=None, float_dg=False, bp_terms=None):
        self.seq, self.macrostates, self.terms, self.motif = seq, dict(macrostates), list(terms), motif
        self.contexts, self.float_dg, self.bp_terms = list(contexts or []), bool(float_dg), list(bp_terms or [])
        names = list(self.macrostates)
        k = self._keep = {}
        k["macro"] = (C.c_char_p * max(len(names), 1))(*([_b(self.macrostates[m]) for m in names] or [b""]))
        T = self.terms
        k["tm"] = (C.c_int * max(len(T), 1))(*([names.index(t[0]) for t in T] or [0]))
        k["tc"] = (C.c_int * max(len(T), 1))(*([1 if t[1] == "holo" else 0 for t in T] or [0]))
        k["tf"] = (C.c_int * max(len(T), 1))(*([1 if t[2] else 0 for t in T] or [0]))
        k["tw"] = (C.c_double * max(len(T), 1))(*([float(t[3]) for t in T] or [0.0]))
        ctx = self.contexts
        k["cb"] = (C.c_char_p * max(len(ctx), 1))(*([_b(c[0]) for c in ctx] or [b""]))
        k["ca"] = (C.c_char_p * max(len(ctx), 1))(*([_b(c[1]) for c in ctx] or [b""]))
        m = _Motif(_b(motif[0]), _b(motif[1]), float(motif[2])) if motif else _Motif(None, None, 0.0)
        bp = (_BpTerm * max(len(self.bp_terms), 1))()
        k["bp"], k["bp_pos"] = bp, []
        for q, t in enumerate(self.bp_terms):
            kind, cond, fav, w = t[0], t[1], t[2], t[3]
            bp[q].kind, bp[q].cond, bp[q].favorable, bp[q].weight = (0 if kind == "pair" else 1), (1 if cond == "holo" else 0), int(bool(fav)), float(w)
            if kind == "pair":
                bp[q].i, bp[q].j, bp[q].n_pos, bp[q].pos = int(t[4]), int(t[5]), 0, None
            else:
                arr = (C.c_int * len(t[4]))(*[int(x) for x in t[4]])
                k["bp_pos"].append(arr)
                bp[q].i, bp[q].j, bp[q].n_pos, bp[q].pos = 0, 0, len(t[4]), arr
        self.c = _Problem(len(seq), _b(seq), len(names), k["macro"], len(T), k["tm"], k["tc"], k["tf"], k["tw"], m,
                          len(ctx), k["cb"], k["ca"], int(self.float_dg), len(self.bp_terms), bp)
        self.n = len(seq)
        self.n_term_values = max(len(ctx), 1) * (len(T) + len(self.bp_terms))

    def as_oracle_dict(self):
        return dict(seq=self.seq, macrostates=self.macrostates, terms=self.terms, motif=self.motif, contexts=self.contexts, bp_terms=self.bp_terms)


def library_path():
    # ADDAPT_B200_LIB: an instrumented build of the same engine (tools/variant_build.sh), for tuning runs only
    return os.environ.get("ADDAPT_B200_LIB") or os.path.join(HERE, "libaddapt_b200.so")


_LIB = None


def load_library():
    """Load libaddapt_b200.so (built in-tree by `make` / __graft_entry__.build()).  Raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise EngineError("libaddapt_b200.so is not built (run `make` or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(path)
    vp, ci, cd, cs = C.c_void_p, C.c_int, C.c_double, C.c_char_p
    L.adp_engine_create.argtypes = [ci, cs, C.POINTER(vp)]
    L.adp_engine_destroy.argtypes = [vp]
    L.adp_engine_destroy.restype = None
    L.adp_last_error.argtypes = [vp]
    L.adp_last_error.restype = cs
    L.adp_engine_set_stream.argtypes = [vp, vp]
    L.adp_engine_set_async.argtypes = [vp, ci]
    L.adp_engine_sync.argtypes = [vp]
    L.adp_engine_set_general_only.argtypes = [vp, ci]
    L.adp_kT_kcal.argtypes = [vp]
    L.adp_kT_kcal.restype = cd
    if hasattr(L, "adp_param_provenance"):        # absent only from older tuning builds selected with ADDAPT_B200_LIB
        L.adp_param_provenance.argtypes = [vp]
        L.adp_param_provenance.restype = cs
        L.adp_param_file.argtypes = [vp]
        L.adp_param_file.restype = cs
    L.adp_device_count.restype = ci
    L.adp_launch_count.argtypes = [vp]
    L.adp_launch_count.restype = C.c_uint64
    L.adp_eval_structure.argtypes = [vp, cs, cs, C.POINTER(C.c_int32)]
    L.adp_motif_contribution.argtypes = [vp, C.POINTER(_Motif), C.POINTER(C.c_int32)]
    L.adp_mfe_batch.argtypes = [vp, ci, ci, vp, cs, C.POINTER(_Motif), vp, vp]
    L.adp_pf_batch.argtypes = [vp, ci, ci, vp, cs, C.POINTER(_Motif), vp, vp]
    L.adp_fold_eval_batch.argtypes = [vp, ci, ci, vp, C.POINTER(_Motif), ci, vp, vp, vp, vp, vp]
    L.adp_score_batch.argtypes = [vp, C.POINTER(_Problem), ci, vp, vp, vp]
    L.adp_mc_run.argtypes = [vp, C.POINTER(_Problem), ci, C.c_uint64, C.c_uint64, ci, ci, C.POINTER(cd), vp, vp, vp, vp, vp, vp]
    L.adp_measure_dfma_rate.argtypes = [vp, C.POINTER(cd)]
    if hasattr(L, "adp_mc_run_trace"):            # absent only from older tuning builds selected with ADDAPT_B200_LIB
        L.adp_mc_trace_width.argtypes = [C.POINTER(_Problem)]
        L.adp_mc_run_trace.argtypes = [vp, C.POINTER(_Problem), ci, C.c_uint64, C.c_uint64, ci, ci, C.POINTER(cd), vp, vp, vp, vp, vp, vp, ci, vp, vp]
        L.adp_mutate_batch.argtypes = [vp, C.POINTER(_Problem), ci, vp, vp, vp, vp, vp]
        L.adp_top_k.argtypes = [vp, ci, vp, ci, vp, vp]
    _LIB = L
    return L


def _ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, int):
        return C.c_void_p(a)
    raise TypeError(type(a))


def seqs_to_array(seqs):
    """list of equal-length strings (or a (B,n) uint8 array) -> contiguous (B,n) uint8 array."""
    if isinstance(seqs, np.ndarray):
        return np.ascontiguousarray(seqs, dtype=np.uint8)
    n = len(seqs[0])
    assert all(len(s) == n for s in seqs), "all sequences of a batch must have the same length"
    return np.frombuffer(b"".join(_b(s) for s in seqs), dtype=np.uint8).reshape(len(seqs), n).copy()


def tri_index(n, i, j):
    """index of pair (i,j), 0-based, in the row-major upper triangle (diagonal included) the engine writes."""
    if i > j:
        i, j = j, i
    return i * n - i * (i - 1) // 2 + (j - i)


class Engine:
    """One engine per GPU.  Thin wrapper: every method is one C-ABI call."""

    def __init__(self, device=0, par_file=None):
        self.lib = load_library()
        h = C.c_void_p()
        # explicit argument > $ADDAPT_B200_PAR > the reconstructed file shipped in params/ (the engine warns on stderr about the
        # latter and reports it through param_provenance: its int11 / int21 / int22 tables are rule-generated)
        par_file = par_file or os.environ.get("ADDAPT_B200_PAR") or DEFAULT_PAR
        rc = self.lib.adp_engine_create(device, _b(par_file), C.byref(h))
        if rc != 0:
            raise EngineError("adp_engine_create failed (%d): %s" % (rc, (self.lib.adp_last_error(None) or b"").decode()))
        self.h = h
        self.kT = self.lib.adp_kT_kcal(self.h)
        has = hasattr(self.lib, "adp_param_provenance")
        self.param_provenance = (self.lib.adp_param_provenance(self.h) or b"").decode() if has else "unknown"
        self.param_file = (self.lib.adp_param_file(self.h) or b"").decode() if has else par_file

    def close(self):
        if getattr(self, "h", None):
            self.lib.adp_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EngineError("engine call failed (%d): %s" % (rc, (self.lib.adp_last_error(self.h) or b"").decode()))

    def _motif(self, motif):
        if motif is None:
            return None, None
        m = _Motif(_b(motif[0]), _b(motif[1]), float(motif[2]))
        return C.byref(m), m

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.adp_engine_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_async(self, flag):
        self._check(self.lib.adp_engine_set_async(self.h, int(flag)))

    def set_general_only(self, flag):
        """route folds through the general kernels instead of the fast ACGU-only ones (parity tests)"""
        self._check(self.lib.adp_engine_set_general_only(self.h, int(flag)))

    def sync(self):
        self._check(self.lib.adp_engine_sync(self.h))

    def launch_count(self):
        return int(self.lib.adp_launch_count(self.h))

    def dfma_rate(self):
        r = C.c_double()
        self._check(self.lib.adp_measure_dfma_rate(self.h, C.byref(r)))
        return r.value

    def eval_structure(self, seq, structure):
        e = C.c_int32()
        self._check(self.lib.adp_eval_structure(self.h, _b(seq), _b(structure), C.byref(e)))
        return e.value

    def motif_contribution(self, motif):
        ref, keep = self._motif(motif)
        c = C.c_int32()
        self._check(self.lib.adp_motif_contribution(self.h, ref, C.byref(c)))
        return c.value

    # ---- batched folds (host numpy in / out) ------------------------------------------------
    def mfe(self, seqs, cst=None, motif=None, structures=True):
        a = seqs_to_array(seqs)
        B, n = a.shape
        e = np.zeros(B, dtype=np.int32)
        st = np.zeros((B, n), dtype=np.uint8) if structures else None
        ref, keep = self._motif(motif)
        self._check(self.lib.adp_mfe_batch(self.h, B, n, _ptr(a), _b(cst) if cst else None, ref, _ptr(e), _ptr(st)))
        return (e, [bytes(r).decode() for r in st]) if structures else e

    def pf(self, seqs, cst=None, motif=None, bppm=False):
        a = seqs_to_array(seqs)
        B, n = a.shape
        dg = np.zeros(B)
        bp = np.zeros((B, n * (n + 1) // 2)) if bppm else None
        ref, keep = self._motif(motif)
        self._check(self.lib.adp_pf_batch(self.h, B, n, _ptr(a), _b(cst) if cst else None, ref, _ptr(dg), _ptr(bp)))
        return (dg, bp) if bppm else dg

    def fold_eval(self, seqs, motif, queries=(), full=False):
        a = seqs_to_array(seqs)
        B, n = a.shape
        qi = np.array([q[0] for q in queries], dtype=np.int32)
        qj = np.array([q[1] for q in queries], dtype=np.int32)
        dg = np.zeros((B, 2))
        pr = np.zeros((B, 2, len(queries))) if len(queries) else None
        bp = np.zeros((B, 2, n * (n + 1) // 2)) if full else None
        ref, keep = self._motif(motif)
        self._check(self.lib.adp_fold_eval_batch(self.h, B, n, _ptr(a), ref, len(queries), _ptr(qi) if len(queries) else None,
                                                 _ptr(qj) if len(queries) else None, _ptr(dg), _ptr(pr), _ptr(bp)))
        return dg, pr, bp

    def fold_eval_raw(self, B, n, seqs, motif, n_query, qi, qj, dg, probs, bppm=None):
        """Pointer-level call for bench.py: every array may be a torch tensor (device or pinned host)."""
        ref, keep = self._motif(motif)
        self._check(self.lib.adp_fold_eval_batch(self.h, B, n, _ptr(seqs), ref, n_query, _ptr(qi), _ptr(qj), _ptr(dg), _ptr(probs), _ptr(bppm)))

    def mfe_raw(self, B, n, seqs, e, structs=None, cst=None, motif=None):
        ref, keep = self._motif(motif)
        self._check(self.lib.adp_mfe_batch(self.h, B, n, _ptr(seqs), _b(cst) if cst else None, ref, _ptr(e), _ptr(structs)))

    def pf_raw(self, B, n, seqs, dg, bppm=None, cst=None, motif=None):
        ref, keep = self._motif(motif)
        self._check(self.lib.adp_pf_batch(self.h, B, n, _ptr(seqs), _b(cst) if cst else None, ref, _ptr(dg), _ptr(bppm)))

    # ---- design problem --------------------------------------------------------------------
    def score(self, problem, seqs, want_terms=False):
        a = seqs_to_array(seqs)
        B, n = a.shape
        assert n == problem.n
        out = np.zeros(B)
        terms = np.zeros((B, problem.n_term_values)) if want_terms else None
        self._check(self.lib.adp_score_batch(self.h, C.byref(problem.c), B, _ptr(a), _ptr(out), _ptr(terms)))
        return (out, terms) if want_terms else out

    def score_raw(self, problem, B, seqs, score, terms=None):
        self._check(self.lib.adp_score_batch(self.h, C.byref(problem.c), B, _ptr(seqs), _ptr(score), _ptr(terms)))

    def top_k(self, scores, k):
        """indices and values of the k best scores (largest first, ties by lower index, NaN last), selected on the device"""
        s = np.ascontiguousarray(scores, dtype=np.float64)
        k = min(int(k), s.shape[0])
        idx, top = np.zeros(k, dtype=np.int32), np.zeros(k)
        self._check(self.lib.adp_top_k(self.h, s.shape[0], _ptr(s), k, _ptr(idx), _ptr(top)))
        return idx, top

    def top_k_raw(self, n, scores, k, top_index, top_score=None):
        self._check(self.lib.adp_top_k(self.h, n, _ptr(scores), k, _ptr(top_index), _ptr(top_score)))

    def mutate(self, problem, seqs, positions, nucleotides):
        """≙ mutate_recursively for every (sequence, position, nucleotide) triple, applied on the device; returns the mutated
        sequences and the can_be_freely_mutated flag of every position (reference src/sampling.cc:177-193, 205-282)."""
        a = seqs_to_array(seqs)
        B, n = a.shape
        assert n == problem.n and len(positions) == B and len(nucleotides) == B
        pos = np.ascontiguousarray(positions, dtype=np.int32)
        nts = np.frombuffer("".join(nucleotides).encode(), dtype=np.uint8).copy()
        out = np.zeros((B, n), dtype=np.uint8)
        free = np.zeros(B, dtype=np.int32)
        self._check(self.lib.adp_mutate_batch(self.h, C.byref(problem.c), B, _ptr(a), _ptr(pos), _ptr(nts), _ptr(out), _ptr(free)))
        return [bytes(r).decode() for r in out], [bool(x) for x in free]

    def mc_run(self, problem, n_traj, seed, n_steps, thermostat=("annealing", 5.0, 0.0, 300), traj_id0=0, trace_interval=0):
        n = problem.n
        if thermostat[0] == "fixed":
            kind, ta = 0, (C.c_double * 3)(thermostat[1], 0, 0)
        elif thermostat[0] == "auto":            # (target acceptance rate, training period, initial temperature)
            kind, ta = 2, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        else:
            kind, ta = 1, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        fs = np.zeros((n_traj, n), dtype=np.uint8)
        bs = np.zeros((n_traj, n), dtype=np.uint8)
        fsc, bsc = np.zeros(n_traj), np.zeros(n_traj)
        bst = np.zeros(n_traj, dtype=np.int32)
        oc = np.zeros((n_traj, 4), dtype=np.uint32)
        tv = tq = None
        if trace_interval > 0 and n_steps > 0:
            cap, width = (n_steps + trace_interval - 1) // trace_interval, self.lib.adp_mc_trace_width(C.byref(problem.c))
            tv = np.zeros((n_traj, cap, width))
            tq = np.zeros((n_traj, cap, 2, n), dtype=np.uint8)
        self._check(self.lib.adp_mc_run_trace(self.h, C.byref(problem.c), n_traj, seed, traj_id0, n_steps, kind, ta,
                                              _ptr(fs), _ptr(fsc), _ptr(bs), _ptr(bsc), _ptr(bst), _ptr(oc), int(trace_interval), _ptr(tv), _ptr(tq)))
        r = dict(final_seq=[bytes(r).decode() for r in fs], final_score=fsc, best_seq=[bytes(r).decode() for r in bs],
                 best_score=bsc, best_step=bst, outcomes=oc)
        if tv is not None:
            r["trace_values"], r["trace_seqs"] = tv, tq      # see adp_mc_run_trace (include/addapt_b200.h) and addapt_b200/trajectory.py
        return r

    def mc_run_raw(self, problem, n_traj, seed, traj_id0, n_steps, thermostat, final_seq=None, final_score=None,
                   best_seq=None, best_score=None, best_step=None, outcomes=None):
        if thermostat[0] == "fixed":
            kind, ta = 0, (C.c_double * 3)(thermostat[1], 0, 0)
        elif thermostat[0] == "auto":            # (target acceptance rate, training period, initial temperature)
            kind, ta = 2, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        else:
            kind, ta = 1, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        self._check(self.lib.adp_mc_run(self.h, C.byref(problem.c), n_traj, seed, traj_id0, n_steps, kind, ta, _ptr(final_seq),
                                        _ptr(final_score), _ptr(best_seq), _ptr(best_score), _ptr(best_step), _ptr(outcomes)))
