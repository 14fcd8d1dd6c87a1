"""ctypes binding of tests/emu/libadp_emu.so -- TEST INFRASTRUCTURE (see tests/emu/emu.cc)."""
import ctypes as C
import os

import numpy as np

from addapt_b200.capi import DEFAULT_PAR, Problem, _Problem, _b

HERE = os.path.dirname(os.path.abspath(__file__))


class Emu:
    def __init__(self, par=None):
        # ADP_EMU_LIB: an emulation build with experiment switches (e.g. -DADP_CORNER_TABLE), for checking them without a GPU
        L = self.lib = C.CDLL(os.environ.get("ADP_EMU_LIB") or os.path.join(HERE, "emu", "libadp_emu.so"))
        L.emu_model_load.restype = C.c_void_p
        L.emu_model_load.argtypes = [C.c_char_p]
        L.emu_error.restype = C.c_char_p
        L.emu_kT.restype = C.c_double
        L.emu_kT.argtypes = [C.c_void_p]
        L.emu_eval.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
        L.emu_motif_contrib.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double, C.POINTER(C.c_int)]
        L.emu_pf.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.c_int,
                             C.POINTER(C.c_double), C.c_void_p]
        L.emu_pf2.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.c_int,
                              C.POINTER(C.c_double), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_mfe.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double,
                              C.POINTER(C.c_int32), C.c_char_p]
        L.emu_mfe2.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double,
                               C.POINTER(C.c_int32), C.c_char_p, C.c_int]
        L.emu_score.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_mc.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double),
                             C.c_char_p, C.POINTER(C.c_double), C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int32),
                             C.POINTER(C.c_uint32), C.c_int]
        self.model = L.emu_model_load(_b(par or DEFAULT_PAR))
        if not self.model:
            raise RuntimeError(L.emu_error().decode())
        self.kT = L.emu_kT(self.model)

    def smem_bytes(self, kind, n):
        """shared memory per construct as the launch planners size it: kind 0 fast PF, 1 fast PF in the design kernels, 2 fast MFE"""
        self.lib.emu_smem_bytes.restype = C.c_size_t
        self.lib.emu_smem_bytes.argtypes = [C.c_int, C.c_int]
        return int(self.lib.emu_smem_bytes(kind, n))

    def _err(self):
        return ValueError(self.lib.emu_error().decode())

    def eval_structure(self, seq, db):
        e = C.c_int()
        if self.lib.emu_eval(self.model, _b(seq), _b(db), C.byref(e)):
            raise self._err()
        return e.value

    def motif_contribution(self, motif):
        c = C.c_int()
        if self.lib.emu_motif_contrib(self.model, _b(motif[0]), _b(motif[1]), motif[2], C.byref(c)):
            raise self._err()
        return c.value

    def pf(self, seq, cst=None, motif=None, bppm=False, fast=False, queries=None):
        """fast=True runs pf_fast.cuh (raises NotImplementedError when it declines the sequence); queries = list of
        0-based (i, j) pairs answered like adp_fold_eval_batch does (returned as a third value)."""
        n = len(seq)
        dg = C.c_double()
        bp = np.zeros(n * (n + 1) // 2)
        nq = len(queries) if queries else 0
        qi = np.array([q[0] for q in queries or []], dtype=np.int32)
        qj = np.array([q[1] for q in queries or []], dtype=np.int32)
        qp = np.full(max(nq, 1), -1.0)
        rc = self.lib.emu_pf2(self.model, n, _b(seq), _b(cst) if cst else None, _b(motif[0]) if motif else None,
                              _b(motif[1]) if motif else None, motif[2] if motif else 0.0, 1 if bppm else 0, C.byref(dg),
                              bp.ctypes.data, 1 if fast else 0, nq, qi.ctypes.data, qj.ctypes.data, qp.ctypes.data)
        if rc == 1:
            raise NotImplementedError("fast path declined the sequence")
        if rc:
            raise self._err()
        if queries is not None:
            return dg.value, bp, qp[:nq]
        return (dg.value, bp) if bppm else dg.value

    def mfe(self, seq, cst=None, motif=None, fast=False):
        n = len(seq)
        e = C.c_int32()
        st = C.create_string_buffer(n + 1)
        rc = self.lib.emu_mfe2(self.model, n, _b(seq), _b(cst) if cst else None, _b(motif[0]) if motif else None,
                               _b(motif[1]) if motif else None, motif[2] if motif else 0.0, C.byref(e), st, 1 if fast else 0)
        if rc == 1:
            raise NotImplementedError("fast path declined the sequence")
        if rc:
            raise self._err()
        return e.value, st.raw[:n].decode()

    def score(self, problem: Problem, seqs, want_terms=False, fast=False):
        B = len(seqs)
        flat = b"".join(_b(s) for s in seqs)
        out = np.zeros(B)
        terms = np.zeros((B, problem.n_term_values))
        if self.lib.emu_score(self.model, C.byref(problem.c), B, flat, out.ctypes.data, terms.ctypes.data, 1 if fast else 0):
            raise self._err()
        return (out, terms) if want_terms else out

    def mutate(self, problem: Problem, seq, pos, nucleotide):
        """one (sequence, position, nucleotide) triple through the table compiler and the device closure code; returns
        (mutated sequence, can_be_freely_mutated)"""
        n = problem.n
        self.lib.emu_mutate.argtypes = [C.POINTER(_Problem), C.c_char_p, C.c_int, C.c_char, C.c_char_p, C.POINTER(C.c_int)]
        out, free = C.create_string_buffer(n + 1), C.c_int()
        if self.lib.emu_mutate(C.byref(problem.c), _b(seq), pos, _b(nucleotide), out, C.byref(free)):
            raise self._err()
        return out.raw[:n].decode(), bool(free.value)

    def mc_run(self, problem: Problem, seed, traj_id, n_steps, thermostat=("annealing", 5.0, 0.0, 300), fast=False, trace_interval=0):
        n = problem.n
        if thermostat[0] == "fixed":
            kind, ta = 0, (C.c_double * 3)(thermostat[1], 0, 0)
        elif thermostat[0] == "auto":            # (target acceptance rate, training period, initial temperature)
            kind, ta = 2, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        else:
            kind, ta = 1, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        fs, bs = C.create_string_buffer(n + 1), C.create_string_buffer(n + 1)
        fsc, bsc, bst = C.c_double(), C.c_double(), C.c_int32()
        oc = (C.c_uint32 * 4)()
        tv = tq = None
        if trace_interval > 0 and n_steps > 0:
            cap = (n_steps + trace_interval - 1) // trace_interval
            tv, tq = np.zeros((cap, 8 + problem.n_term_values)), np.zeros((cap, 2, n), dtype=np.uint8)
        self.lib.emu_mc_trace.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double),
                                          C.c_char_p, C.POINTER(C.c_double), C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                          C.POINTER(C.c_uint32), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        if self.lib.emu_mc_trace(self.model, C.byref(problem.c), seed, traj_id, n_steps, kind, ta, fs, C.byref(fsc), bs, C.byref(bsc),
                                 C.byref(bst), oc, 1 if fast else 0, int(trace_interval), tv.ctypes.data if tv is not None else None,
                                 tq.ctypes.data if tq is not None else None):
            raise self._err()
        return dict(final_seq=fs.raw[:n].decode(), final_score=fsc.value, best_seq=bs.raw[:n].decode(), best_score=bsc.value,
                    best_step=bst.value, outcomes=list(oc), trace_values=tv, trace_seqs=tq)

    # ---- the emulated CTA with all of its threads (tests/emu/simt.hh, emu_simt.cc) -------------------------------------------
    POLICIES = {"random": 0, "ascending": 1, "descending": 2, "random_warp": 3}

    def _simt_err(self):
        self.lib.emu_simt_error.restype = C.c_char_p
        return RuntimeError("simt emulation: " + self.lib.emu_simt_error().decode())

    @staticmethod
    def _strings(items):
        arr = (C.c_char_p * len(items))(*[_b(s) if s else None for s in items])
        return arr

    def simt_pf(self, seqs, csts=None, motif=None, use_motif=None, outside=None, path="fast", threads=256, group=1, policy="random",
                seed=0):
        """several partition-function jobs of one length back to back on one emulated CTA.  path: general | fast | fast_qm_ext |
        fast_mem_levels.  -> dict(dG, bppm, declined, stats=(CTA barriers, warp collectives, context switches))"""
        J, n = len(seqs), len(seqs[0])
        csts = csts or [None] * J
        use_motif = np.array(use_motif if use_motif is not None else [1 if motif else 0] * J, dtype=np.int32)
        outside = np.array(outside if outside is not None else [1] * J, dtype=np.int32)
        dg, bp = np.zeros(J), np.zeros((J, n * (n + 1) // 2))
        declined, stats = np.zeros(J, dtype=np.int32), np.zeros(3, dtype=np.int64)
        self.lib.emu_simt_pf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_char_p,
                                         C.c_char_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        mp = C.c_uint64()
        rc = self.lib.emu_simt_pf(self.model, J, n, b"".join(_b(s) for s in seqs), self._strings(csts), use_motif.ctypes.data,
                                  outside.ctypes.data, _b(motif[0]) if motif else None, _b(motif[1]) if motif else None,
                                  motif[2] if motif else 0.0, dg.ctypes.data, bp.ctypes.data, declined.ctypes.data,
                                  {"general": 0, "fast": 1, "fast_qm_ext": 2, "fast_mem_levels": 3}[path], threads, group,
                                  self.POLICIES[policy], seed, stats.ctypes.data, C.byref(mp))
        if rc:
            raise self._simt_err()
        # matrices: fingerprint of the matrix block as the LAST job left it (every cell, not only the outputs)
        return dict(dG=dg, bppm=bp, declined=declined, stats=tuple(int(v) for v in stats), matrices=mp.value)

    def simt_mfe(self, seqs, csts=None, motif=None, use_motif=None, path="fast", threads=256, group=1, policy="random", seed=0):
        J, n = len(seqs), len(seqs[0])
        csts = csts or [None] * J
        use_motif = np.array(use_motif if use_motif is not None else [1 if motif else 0] * J, dtype=np.int32)
        e, st = np.zeros(J, dtype=np.int32), np.zeros((J, n), dtype=np.uint8)
        declined, stats = np.zeros(J, dtype=np.int32), np.zeros(3, dtype=np.int64)
        self.lib.emu_simt_mfe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_char_p, C.c_char_p,
                                          C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                          C.c_void_p, C.POINTER(C.c_uint64)]
        mp = C.c_uint64()
        rc = self.lib.emu_simt_mfe(self.model, J, n, b"".join(_b(s) for s in seqs), self._strings(csts), use_motif.ctypes.data,
                                   _b(motif[0]) if motif else None, _b(motif[1]) if motif else None, motif[2] if motif else 0.0,
                                   e.ctypes.data, st.ctypes.data, declined.ctypes.data, {"general": 0, "fast": 1, "fast_mem_levels": 3}[path],
                                   threads, group, self.POLICIES[policy], seed, stats.ctypes.data, C.byref(mp))
        if rc:
            raise self._simt_err()
        return dict(e=e, structures=[bytes(r).decode("latin-1") for r in st], declined=declined, stats=tuple(int(v) for v in stats),
                    matrices=mp.value)

    def simt_score(self, problem: Problem, seqs, fast=True, qm_global=False, threads=256, policy="random", seed=0):
        B = len(seqs)
        out, terms, stats = np.zeros(B), np.zeros((B, problem.n_term_values)), np.zeros(3, dtype=np.int64)
        self.lib.emu_simt_score.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        rc = self.lib.emu_simt_score(self.model, C.byref(problem.c), B, b"".join(_b(s) for s in seqs), out.ctypes.data, terms.ctypes.data,
                                     1 if fast else 0, 1 if qm_global else 0, threads, self.POLICIES[policy], seed, stats.ctypes.data)
        if rc:
            raise self._simt_err()
        return out, terms, tuple(int(v) for v in stats)

    def simt_mc(self, problem: Problem, seed, traj_id0, n_traj, n_steps, thermostat=("annealing", 5.0, 0.0, 300), fast=True, threads=256,
                policy="random", sched_seed=0):
        n = problem.n
        if thermostat[0] == "fixed":
            kind, ta = 0, (C.c_double * 3)(thermostat[1], 0, 0)
        elif thermostat[0] == "auto":
            kind, ta = 2, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        else:
            kind, ta = 1, (C.c_double * 3)(thermostat[1], thermostat[2], thermostat[3])
        fs, bs = np.zeros((n_traj, n), dtype=np.uint8), np.zeros((n_traj, n), dtype=np.uint8)
        fsc, bsc, bst = np.zeros(n_traj), np.zeros(n_traj), np.zeros(n_traj, dtype=np.int32)
        oc, stats = np.zeros((n_traj, 4), dtype=np.uint32), np.zeros(3, dtype=np.int64)
        self.lib.emu_simt_mc.argtypes = [C.c_void_p, C.POINTER(_Problem), C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        rc = self.lib.emu_simt_mc(self.model, C.byref(problem.c), seed, traj_id0, n_traj, n_steps, kind, ta, fs.ctypes.data, fsc.ctypes.data,
                                  bs.ctypes.data, bsc.ctypes.data, bst.ctypes.data, oc.ctypes.data, 1 if fast else 0, threads,
                                  self.POLICIES[policy], sched_seed, stats.ctypes.data)
        if rc:
            raise self._simt_err()
        return [dict(final_seq=bytes(fs[k]).decode(), final_score=float(fsc[k]), best_seq=bytes(bs[k]).decode(), best_score=float(bsc[k]),
                     best_step=int(bst[k]), outcomes=[int(v) for v in oc[k]]) for k in range(n_traj)]

    def simt_selftest_missing_barrier(self, threads, policy, seed, with_barrier):
        s = C.c_long()
        self.lib.emu_simt_selftest_missing_barrier.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_long)]
        if self.lib.emu_simt_selftest_missing_barrier(threads, self.POLICIES[policy], seed, 1 if with_barrier else 0, C.byref(s)):
            raise self._simt_err()
        return s.value

    def simt_selftest_divergent_collective(self, threads, policy, seed):
        self.lib.emu_simt_selftest_divergent_collective.argtypes = [C.c_int, C.c_int, C.c_uint64]
        if self.lib.emu_simt_selftest_divergent_collective(threads, self.POLICIES[policy], seed):
            raise self._simt_err()

    def simt_drop_barrier(self, k):
        """fault injection: every thread of the following simt_* runs returns from its k-th CTA barrier at once (-1: off)"""
        self.lib.emu_simt_drop_barrier.argtypes = [C.c_long]
        self.lib.emu_simt_drop_barrier(k)
