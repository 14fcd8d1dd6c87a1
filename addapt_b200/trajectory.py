"""On-disk trajectory format of the reference (TsvTrajectoryReporter, reference src/sampling.cc:430-489) for runs of the GPU
Monte-Carlo kernel: adp_mc_run_trace returns per-step records (include/addapt_b200.h); write_tsv lays one trajectory out with the
reference's header line, column names, column order and number formatting, so that the reference's own tools
(tools/utils.py:6-51: pandas.read_table(comment='#'), load_initial_seq, pick_best_scores over 'current_score') read it unchanged.
"""
import math

OUTCOMES = ("REJECT", "ACCEPT_WORSENED", "ACCEPT_UNCHANGED", "ACCEPT_IMPROVED")     # reference src/sampling.cc:501-510
MOVE_NAME = "UnbiasedMutation"                                                      # reference include/sampling.hh:139


def fmt(x):
    """a double the way `std::ostream << double` prints it with default flags (6 significant digits, like printf %g)"""
    x = float(x)
    if math.isnan(x):
        return "-nan" if math.copysign(1.0, x) < 0 else "nan"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    return "%g" % x


def term_names(problem):
    """names of the score-table rows, in the engine's order [context][macrostate terms, bp terms]: MacrostateProbTerm's
    "apo|holo: [not ]<macrostate>" (reference src/scoring.cc:213-231), prefixed "<context>: " when contexts are named (:147-150)"""
    rows = ["%s: %s%s" % (cond, "" if fav else "not ", macro) for macro, cond, fav, _w in problem.terms]
    for t in getattr(problem, "bp_terms", []):
        what = "pair %d-%d" % (t[4], t[5]) if t[0] == "pair" else "paired density %s" % ",".join(str(p) for p in t[4])
        rows.append("%s: %s%s" % (t[1], "" if t[2] else "not ", what))
    return rows


def term_weights(problem):
    return [float(t[3]) for t in problem.terms] + [float(t[3]) for t in getattr(problem, "bp_terms", [])]


def rows(problem, values, seqs, num_steps, context_names=None):
    """one list of column strings per record of ONE trajectory (values: [cap, width], seqs: [cap, 2, n])"""
    names, weights = term_names(problem), term_weights(problem)
    n_ctx = max(len(problem.contexts), 1)
    out = []
    for v, q in zip(values, seqs):
        r = ["%d" % int(v[0]), "%d" % num_steps, fmt(v[1]), fmt(v[2])]
        for x in range(n_ctx):
            for t in range(len(names)):
                r += [fmt(weights[t]), fmt(v[8 + x * len(names) + t])]
        r += [fmt(v[3]), fmt(v[4]), fmt(v[5]), fmt(v[6]), MOVE_NAME, OUTCOMES[int(v[7])], bytes(q[0]).decode(), bytes(q[1]).decode()]
        out.append(r)
    return out


def header(problem, context_names=None):
    names = term_names(problem)
    n_ctx = max(len(problem.contexts), 1)
    ctx = list(context_names) if context_names else ([""] if not problem.contexts else ["%d" % x for x in range(n_ctx)])
    cols = ["step", "num_steps", "current_score", "proposed_score"]
    for x in range(n_ctx):
        prefix = ctx[x] + ": " if problem.contexts else ""
        for nm in names:
            cols += ["term_weight[%s%s]" % (prefix, nm), "term_value[%s%s]" % (prefix, nm)]
    return cols + ["score_diff", "temperature", "metropolis_criterion", "random_threshold", "move", "outcome", "current_seq", "proposed_seq"]


def write_tsv(path, problem, values, seqs, num_steps, context_names=None):
    """values / seqs: the records of ONE trajectory (trace_values[t], trace_seqs[t] of Engine.mc_run(..., trace_interval=k)).
    Every line ends with a tab before the newline, exactly like the reference's stream of `<< "\\t"`."""
    with open(path, "w") as f:
        f.write("#\tinitial_seq\t%s\n" % problem.seq)        # Device::seq() of the start device: mixed case as given
        f.write("\t".join(header(problem, context_names)) + "\t\n")
        for r in rows(problem, values, seqs, num_steps, context_names):
            f.write("\t".join(r) + "\t\n")


# ---- reading trajectory files back: what the reference's tools/utils.py does with them (pandas-free) -----------------------------

def load_initial_seq(tsv_path):
    """the '# initial_seq <seq>' line a trajectory file starts with (reference tools/utils.py:12-25)"""
    with open(tsv_path) as f:
        header = f.readline()
    if header.startswith("#"):
        _, label, value = header.rstrip("\n").split("\t")[:3]
        if label == "initial_seq":
            return value
    return None


def load_trajectory(tsv_path):
    """column name -> list of values (floats where the text parses as one), comment lines skipped (reference tools/utils.py:6-7:
    pandas.read_table(path, comment='#'))"""
    cols, data = None, None
    with open(tsv_path) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            cells = line.rstrip("\n").split("\t")
            if cells and cells[-1] == "":
                cells = cells[:-1]                       # every line of the format ends with a tab
            if cols is None:
                cols, data = cells, [[] for _ in cells]
                continue
            for k, c in enumerate(cells[:len(cols)]):
                try:
                    data[k].append(float(c))
                except ValueError:
                    data[k].append(c)
    return dict(zip(cols or [], data or []))


def pick_best_seqs(tsv_paths, window_size):
    """The designs the reference's offline tools would pick from a set of trajectory files (reference tools/utils.py:27-67):
    per file, the peaks of the current_score column (sharding.pick_best_scores: arg-max, blank +-window, while >= 75th percentile),
    each returned as (path, record index, step, score, sequence).  The snapshot's files carry the whole sequence in current_seq
    (the reference's extract_pick still concatenates the 'current_domain*' columns of an older format)."""
    from .sharding import pick_best_scores
    picks = []
    for path in tsv_paths:
        traj = load_trajectory(path)
        if not traj or not traj.get("current_score"):
            continue
        for i in pick_best_scores(traj["current_score"], window_size):
            picks.append((path, i, int(traj["step"][i]), traj["current_score"][i], traj["current_seq"][i]))
    return picks
