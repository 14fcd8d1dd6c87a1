#!/usr/bin/env python3
"""Writes tests/golden/reference_thresholds.json: every expectation the reference's OWN tests hold for the
fold / score / move path, transcribed by hand from /root/reference (read-only, not available on the GPU box)
with the file:line each block comes from.  These are the only "golden vectors" the reference has for this
path: inequality thresholds for ViennaRNA-backed results (tests/test_scoring.cc:83-259), exact tables for
the move and score-term logic (tests/test_sampling.cc:8-112, tests/test_scoring.cc:332-384), and five
energies that appear only in comments (tests/test_scoring.cc:54-55,87,153-154).
Run:  python tests/golden/make_reference_golden.py
"""
import json
import os

G = {}

G["hairpin"] = {   # tests/test_scoring.cc:83-147
    "source": "tests/test_scoring.cc:83-147",
    "seq": "ACGUGAAAACGU",
    "comment_structure": "((((....))))", "comment_energy_kcal": -2.20,          # :87
    "pairs_gt": [[0, 11, 0.70], [1, 10, 0.95], [2, 9, 0.95], [3, 8, 0.95]],      # :95-101
    "other_pairs_lt": 0.1,                                                       # :115
    "oob": [[0, 12], [0, -13]],                                                  # :131-132
    "macrostates": [["...(....)...", ">", 0.95], ["..((....))..", ">", 0.85], [".(((....))).", ">", 0.75],
                    ["((((....))))", ">", 0.65], ["(((......)))", ">", 0.65], ["((........))", ">", 0.65],
                    ["(..........)", ">", 0.65], ["xxxxxxxxxxxx", "<", 0.05], ["xxxx........", "<", 0.05],
                    ["....xxxx....", ">", 0.95], ["........xxxx", "<", 0.05]],   # :135-146
}

G["aptamer"] = {   # tests/test_scoring.cc:149-195
    "source": "tests/test_scoring.cc:149-195",
    "seq": "GAUACCAGCCGAAAGGCCCUUGGCAGC",
    "motif": ["GAUACCAGCCGAAAGGCCCUUGGCAGC", "(...((.(((....)))....))...)", 0.32],   # :45-48
    "comment_apo": ["....((((((....)))...)))....", -6.20], "comment_holo": ["(...((.(((....)))....))...)", -9.22],   # :153-154
    "apo_pairs": [[4, 22], [5, 21], [6, 20], [7, 16], [8, 15], [9, 14]],           # :159-161
    "holo_pairs": [[0, 26], [4, 22], [5, 21], [7, 16], [8, 15], [9, 14]],          # :162-164
    "present_gt": 0.7, "absent_lt": 0.3,                                           # :175-186
    "macrostates": [["apo", "....((((((....)))...)))....", ">", 0.65], ["apo", "(.........................)", "<", 0.05],
                    ["holo", "....((((((....)))...)))....", "<", 0.01], ["holo", "(.........................)", ">", 0.95]],  # :191-194
}

G["rhf6"] = {      # tests/test_scoring.cc:45-63,197-259
    "source": "tests/test_scoring.cc:45-63,197-259",
    "seq": "guuuuagagcuagaaauagcaaguuaaaauaaggcuaguccCuUUUCGCCgauaccagccgaaaggcccuuggcagcGACggcaccgagucggugcuuuuuu",
    "active": "(............................)xx..xxxxx..xxxxxx(...............................).(.............)......",
    "comment_dG_apo": -29.58, "comment_dG_holo": -33.82,                           # :54-55
    "constitutive": [[0, 29, 0.75], [1, 28, 0.80], [2, 27, 0.85], [3, 26, 0.85], [4, 25, 0.85], [5, 24, 0.85], [6, 23, 0.80],
                     [8, 19, 0.95], [9, 18, 0.95], [10, 17, 0.95], [11, 16, 0.95], [81, 95, 0.90], [82, 94, 0.90],
                     [83, 93, 0.90], [84, 92, 0.90], [85, 91, 0.90], [86, 90, 0.75]],   # :201-208
    "apo_pairs": [[33, 73, 0.30], [34, 72, 0.30], [35, 71, 0.30], [36, 70, 0.25], [37, 69, 0.10], [40, 65, 0.35], [41, 64, 0.35],
                  [42, 63, 0.35], [43, 62, 0.35], [44, 61, 0.30], [45, 60, 0.25], [47, 58, 0.35], [48, 57, 0.40]],   # :209-214
    "apo_pairs_holo_lt": 1e-3,                                                     # :239
    "holo_pairs": [[30, 43, 0.55, 0.25], [31, 42, 0.75, 0.35], [32, 41, 0.75, 0.35], [33, 40, 0.75, 0.35], [46, 80, 0.50, 0.05],
                   [47, 79, 0.60, 0.05], [49, 77, 0.85, 0.05], [50, 76, 0.95, 0.05], [54, 72, 0.95, 0.20], [55, 71, 0.95, 0.20],
                   [57, 66, 0.95, 0.50], [58, 65, 0.95, 0.50], [59, 64, 0.95, 0.50]],   # :215-222  [i, j, holo >, apo <]
    "p_apo_active_lt": 7e-5, "p_holo_active_gt": 4e-3,                             # :257-258
}

G["mutate_recursively"] = {   # tests/test_sampling.cc:47-112
    "source": "tests/test_sampling.cc:47-112",
    "table": [["N", ["."], "A", ["A"]],
              ["NN", ["()"], "AG", ["AU", "CG"]],
              ["NNN", ["(.)"], "AGU", ["ANU", "NGN", "ANU"]],
              ["NN", ["()", "()"], "AG", ["AU", "CG"]],
              ["NNN", ["().", ".()"], "AGU", ["AUA", "CGC", "UAU"]],
              ["NNN", ["().", "(.)"], "AGU", ["AUU", "CGG", "AUU"]],
              ["NNNN", ["()..", "(())"], "AGUC", ["AUAU", "CGCG", "UAUA", "GCGC"]],
              ["NNNN", ["(.).", "(())"], "AGUC", ["AAUU", "GGCC", "AAUU", "GGCC"]]],
    "throws": [["Nn", ["()"], 0, "G"], ["N", ["("], 0, "G"], ["N", [")"], 0, "G"]],
}

G["mutable_positions"] = {    # tests/test_sampling.cc:8-45
    "source": "tests/test_sampling.cc:8-45",
    "cases": [{"seq": "UUUuuu", "macrostates": ["(.)(.)"], "can_be_mutated": [1, 1, 1, 0, 0, 0], "can_be_freely_mutated": [1, 1, 0, 0, 0, 0]},
              {"seq": "UUUU", "macrostates": ["().)", "(.))"], "can_be_mutated": [1, 1, 1, 1], "can_be_freely_mutated": [1, 0, 0, 0]}],
}

G["macrostate_prob_term"] = {  # tests/test_scoring.cc:332-384: [name, condition, favorable, apo_p, holo_p, expected = log(x)]
    "source": "tests/test_scoring.cc:332-384",
    "rows": [["apo: not dummy", "apo", 0, 0.2, 0.2, 0.8], ["apo: not dummy", "apo", 0, 0.9, 0.2, 0.1], ["apo: not dummy", "apo", 0, 0.2, 0.9, 0.8],
             ["apo: not dummy", "apo", 0, 0.9, 0.9, 0.1], ["apo: dummy", "apo", 1, 0.2, 0.2, 0.2], ["apo: dummy", "apo", 1, 0.9, 0.2, 0.9],
             ["apo: dummy", "apo", 1, 0.2, 0.9, 0.2], ["apo: dummy", "apo", 1, 0.9, 0.9, 0.9], ["holo: not dummy", "holo", 0, 0.2, 0.2, 0.8],
             ["holo: not dummy", "holo", 0, 0.9, 0.2, 0.8], ["holo: not dummy", "holo", 0, 0.2, 0.9, 0.1], ["holo: not dummy", "holo", 0, 0.9, 0.9, 0.1],
             ["holo: dummy", "holo", 1, 0.2, 0.2, 0.2], ["holo: dummy", "holo", 1, 0.9, 0.2, 0.2], ["holo: dummy", "holo", 1, 0.2, 0.9, 0.9],
             ["holo: dummy", "holo", 1, 0.9, 0.9, 0.9]],
}

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_thresholds.json")
with open(out, "w") as f:
    json.dump(G, f, indent=1)
print("wrote", out)
