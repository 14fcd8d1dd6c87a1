// Host facade of addapt_b200 -- scoring.  Same interface as the reference's include/scoring.hh: the abstract fold
// engine RnaFold (:42-55), ScoreFunction (:97-142), ScoreTerm (:144-171) and MacrostateProbTerm (:175-192).
// Where the reference hard-codes ViennaRnaFold (src/scoring.cc:145-146) this facade asks a FoldFactory, whose
// default builds B200RnaFold objects on the GPU engine (addapt/b200.hh).  There is no CPU fold engine here.
#pragma once
#include <functional>

#include "addapt/model.hh"

namespace addapt {

class ScoreFunction;
using ScoreFunctionPtr = std::shared_ptr<ScoreFunction>;

struct EvaluatedScoreTerm { string name; double weight, term; };
using EvaluatedScoreFunction = std::vector<EvaluatedScoreTerm>;

class ScoreTerm;
using ScoreTermPtr = std::shared_ptr<ScoreTerm>;
using ScoreTermList = std::vector<ScoreTermPtr>;

enum class ConditionEnum { APO, HOLO };
enum class FavorableEnum { NO, YES };

// The interface to RNA secondary-structure predictions (reference include/scoring.hh:42-55).
// One object == one (sequence, condition) pair.
class RnaFold {
 public:
  virtual ~RnaFold() {}
  // Probability that positions a and b pair: 0-based, negatives count from the end, order-free, a == b -> 0,
  // out of range throws std::string (reference src/scoring.cc:37-51).
  virtual double base_pair_prob(int a, int b) const = 0;
  // Probability of the macrostate given as a hard-constraint string over the (contextual) sequence
  // (reference src/scoring.cc:53-71).
  virtual double macrostate_prob(string constraint) const = 0;
};
using RnaFoldPtr = std::shared_ptr<RnaFold>;

// Builds the fold engine for a device and an optional aptamer (nullptr = apo).
using FoldFactory = std::function<RnaFoldPtr(DeviceConstPtr, AptamerConstPtr)>;

class ScoreFunction {
 public:
  ScoreFunction();
  virtual ~ScoreFunction() {}

  double evaluate(DeviceConstPtr) const;
  // Also fills the table of (name, weight, unweighted value) rows.  With contexts the score is the SUM over the
  // contexts in name order and rows are prefixed "<context>: " (reference src/scoring.cc:114-135).
  virtual double evaluate(DeviceConstPtr, EvaluatedScoreFunction&) const;

  void add_term(ScoreTermPtr);
  void operator+=(ScoreTermPtr);
  ScoreTermList terms() const;
  AptamerConstPtr aptamer() const;
  void aptamer(AptamerConstPtr);
  ContextConstPtr context(string name) const;
  void add_context(string name, ContextConstPtr);
  map<string, ContextConstPtr> const& contexts() const;

  // Fold engine used by evaluate_terms (default: the B200 engine; tests inject fakes).
  void fold_factory(FoldFactory);
  FoldFactory fold_factory() const;

 protected:
  double evaluate_terms(DeviceConstPtr, EvaluatedScoreFunction&, string prefix = "") const;

 private:
  ScoreTermList my_terms;
  AptamerConstPtr my_aptamer;
  map<string, ContextConstPtr> my_contexts;
  FoldFactory my_factory;
};

class ScoreTerm {
 public:
  ScoreTerm(string name = "", double weight = 1.0);
  virtual ~ScoreTerm() {}
  // unweighted value of the term for this device in the two conditions
  virtual double evaluate(DeviceConstPtr, RnaFold const& apo, RnaFold const& holo) const = 0;
  string name() const;
  void name(string);
  double weight() const;
  void weight(double);

 private:
  string my_name;
  double my_weight;
};

// log P(macrostate) -- or log(1 - P) when the macrostate is unwanted -- in one condition
// (reference src/scoring.cc:204-259).  Name: "apo: not active", "holo: active", ...
class MacrostateProbTerm : public ScoreTerm {
 public:
  MacrostateProbTerm(string macrostate, ConditionEnum, FavorableEnum = FavorableEnum::YES);
  double evaluate(DeviceConstPtr, RnaFold const& apo, RnaFold const& holo) const;
  string macrostate() const { return my_macrostate; }
  ConditionEnum condition() const { return my_condition; }
  FavorableEnum favorable() const { return my_favorable; }

 private:
  string my_macrostate;
  ConditionEnum my_condition;
  FavorableEnum my_favorable;
};

// Score terms over RnaFold::base_pair_prob (reference include/scoring.hh:48).  The mounted snapshot of the reference ships only
// MacrostateProbTerm, but keeps base_pair_prob on the interface and a fake backend for such terms (tests/test_scoring.cc:12-43);
// these two restate the historic kinds BASELINE.json's north_star names.  Positions are 0-based indices into the RAW device
// (a context shifts them by the length of its 5' flank).  B200ScoreFunction evaluates both on the device (adp_bp_term).

// log P(i pairs with j) -- or log(1 - P) -- in one condition ("conditional pairing").  Name: "holo: pair 47-79", "apo: not pair 47-79"
class PairProbTerm : public ScoreTerm {
 public:
  PairProbTerm(int i, int j, ConditionEnum, FavorableEnum = FavorableEnum::YES);
  double evaluate(DeviceConstPtr, RnaFold const& apo, RnaFold const& holo) const;
  int i() const { return my_i; }
  int j() const { return my_j; }
  ConditionEnum condition() const { return my_condition; }
  FavorableEnum favorable() const { return my_favorable; }

 private:
  int my_i, my_j;
  ConditionEnum my_condition;
  FavorableEnum my_favorable;
};

// log of the mean probability of being paired over a set of positions -- or log(1 - mean) -- in one condition ("bppm density")
class PairedDensityTerm : public ScoreTerm {
 public:
  PairedDensityTerm(vector<int> positions, ConditionEnum, FavorableEnum = FavorableEnum::YES);
  double evaluate(DeviceConstPtr, RnaFold const& apo, RnaFold const& holo) const;
  vector<int> const& positions() const { return my_positions; }
  ConditionEnum condition() const { return my_condition; }
  FavorableEnum favorable() const { return my_favorable; }

 private:
  vector<int> my_positions;
  ConditionEnum my_condition;
  FavorableEnum my_favorable;
};

}  // namespace addapt

namespace std {
ostream& operator<<(ostream&, const addapt::ConditionEnum&);
ostream& operator<<(ostream&, const addapt::FavorableEnum&);
}
