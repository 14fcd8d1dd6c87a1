// Host facade of addapt_b200 -- configuration.  Same entry points as the reference's include/config.hh
// (device_from_yaml, scorefxn_from_yaml, score_term_from_str, thermostat_from_yaml, thermostat_from_str;
// reference src/config.cc:54-175).  yaml-cpp is not available in this image, so the *_from_yaml functions read the
// small YAML subset the reference's schema needs (top-level scalars, one level of maps, inline [a, b] lists).
#pragma once
#include "addapt/model.hh"
#include "addapt/sampling.hh"
#include "addapt/scoring.hh"

namespace addapt {

// sections: `sequence`, `macrostates` (name: constraint)                     reference src/config.cc:54-70
DevicePtr device_from_yaml(vector<string> config_files);
// sections: `objective` {apo, holo}, `aptamer` {sequence, fold, affinity}, optional `contexts` {name: [5', 3']}
// -> a B200ScoreFunction with one MacrostateProbTerm per condition             reference src/config.cc:73-101
ScoreFunctionPtr scorefxn_from_yaml(vector<string> config_files);
// "[not ]<macrostate>"                                                        reference src/config.cc:104-115
ScoreTermPtr score_term_from_str(ConditionEnum, string spec);
// "<apo|holo> [not] pair <i> <j> [x <w>]" | "<apo|holo> [not] density <first>-<last> [x <w>]"  (optional `pair_objective` section)
ScoreTermPtr bp_term_from_str(string spec);
// optional section `thermostat` (default "1")                                 reference src/config.cc:118-122
ThermostatPtr thermostat_from_yaml(vector<string> config_files);
// "<T>" | "<T_hi> to <T_lo> in <N> steps" | "auto [<P>% [<period> [<T0>]]]"   reference src/config.cc:124-175
ThermostatPtr thermostat_from_str(string spec);

}  // namespace addapt
